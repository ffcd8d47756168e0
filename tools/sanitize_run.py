"""Short run of every mode for compute-sanitizer (memcheck / initcheck): NeRF training with the camera optimisation, a classic and a Blender frame with
masks, the image and SDF modes.  compute-sanitizer --tool memcheck python tools/sanitize_run.py
`python tools/sanitize_run.py k19`: the error-map importance sampling and exposure paths only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "blender-ngp_b200"), os.path.join(ROOT, "tests")]
import pyngp
import synthetic
from golden_inputs import procedural_image, sdf_pool

scene = synthetic.make_lego_scene(8, 64, device="cpu", seed=0)
if "k19" in sys.argv[1:]:
    # error-map importance sampling + exposure optimisation only: two windows, so that CDFs are built, drawn from (K1, K6, camera gradient) and rebuilt
    tb = pyngp.Testbed()
    tb.load_training_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
    tr = tb.nerf.training
    tr.sample_focal_plane_proportional_to_error = tr.sample_image_proportional_to_error = True
    tr.optimize_exposure = tr.optimize_extrinsics = True
    tb.train_n(128 + 192 + 3, 1 << 12)
    print("k19 loss", tb.loss, "pmf", tr.get_error_map_pmf(), "window", tr.n_steps_between_error_map_updates)
    sys.exit(0)
tb = pyngp.Testbed()
tb.load_training_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
tb.nerf.training.optimize_extrinsics = True
tb.train_n(40, 1 << 14)
print("nerf loss", tb.loss, "offset", tb.nerf.training.get_camera_offsets(0)[0])
tb.set_nerf_camera_matrix(np.asarray(scene["nerf_c2w"][0])[:3])
print("classic frame", tb.render(96, 64, 2, True).mean())
snap = "/tmp/ngpb_sanitize.msgpack"
tb.save_snapshot(snap)
res = (96, 64)
out_p = pyngp.RenderOutputProperties(res, pyngp.DownsampleInfo.MakeFromMip(res, 0), 1, pyngp.ColorSpace.SRGB, pyngp.TonemapCurve.ACES, 0.0, [0, 0, 0, 0], False)
cam_p = pyngp.RenderCameraProperties(synthetic.nerf_matrix_to_ngp(np.asarray(scene["nerf_c2w"][0])), pyngp.CameraModel.Perspective, scene["fx"] * 96 / 64, 0.0, 0.0, 1.0, None, None)
box = pyngp.BoundingBox([0, 0, 0], [1, 1, 1])
mask = pyngp.Mask3D.Sphere(0.3, np.eye(4), pyngp.MaskMode.Add, 0.05, 1.0)
rq = pyngp.RenderRequest(out_p, cam_p, pyngp.RenderModifiers([]), [pyngp.NerfDescriptor(snap, box, np.eye(4), pyngp.RenderModifiers([mask]), 0.8)], box)
print("blender frame", tb.request_nerf_render_sync(rq).mean())
img = pyngp.Testbed(pyngp.TestbedMode.Image)
img.load_image_data(procedural_image())
img.train_n(20, 1 << 14)
print("image loss", img.loss, "mse", img.compute_image_mse(), "frame", img.render(64, 48, 2, False).mean())
sdf = pyngp.Testbed(pyngp.TestbedMode.Sdf)
sdf.set_unit_cube_pairs(*sdf_pool())
sdf.train_n(20, 1 << 14)
print("sdf loss", sdf.loss)
