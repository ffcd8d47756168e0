"""Kernel timeline of a few training steps through CUPTI (torch.profiler), since nsys is not in the image.
usage: python tools/timeline.py [n_steps] [out.json]  -> prints per-kernel start/duration/stream for the last steps."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp
import synthetic

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "train_n"
scene = synthetic.make_lego_scene(100, 800, device="cuda", as_numpy=True)
tb = pyngp.Testbed()
tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
tb.train_n(530)
rq = None
if mode == "blender":  # one 800 x 800 frame through request_nerf_render_sync (K18), one NeRF from a snapshot
    import math
    import numpy as np
    snap = "/tmp/ngpb_timeline.msgpack"
    tb.save_snapshot(snap)
    cam = synthetic.nerf_matrix_to_ngp(synthetic.hemisphere_cameras(7, seed=3)[2])
    out_p = pyngp.RenderOutputProperties((800, 800), pyngp.DownsampleInfo.MakeFromMip((800, 800), 0), 1, pyngp.ColorSpace.SRGB, pyngp.TonemapCurve.Identity, 0.0, [0, 0, 0, 0], False)
    cam_p = pyngp.RenderCameraProperties(cam, pyngp.CameraModel.Perspective, scene["fx"], 0.0, 0.0, 1.0, None, None)
    box = pyngp.BoundingBox([0, 0, 0], [1, 1, 1])
    rq = pyngp.RenderRequest(out_p, cam_p, pyngp.RenderModifiers([]), [pyngp.NerfDescriptor(snap, box, np.eye(4), pyngp.RenderModifiers([]), 1.0)], box)
    tb.request_nerf_render_sync(rq)
if mode == "render":
    import math
    tb.camera_matrix = synthetic.nerf_matrix_to_ngp(synthetic.hemisphere_cameras(7, seed=3)[2])
    tb.fov_axis = 0; tb.fov = math.degrees(synthetic.CAMERA_ANGLE_X)
    tb.render(800, 800, 1, True)  # warm-up: workspace allocation
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    if mode == "train_n":
        tb.train_n(n_steps)
    elif mode == "blender":
        tb.request_nerf_render_sync(rq)
    elif mode == "render":  # one classic 800 x 800 frame (K17)
        import math
        tb.camera_matrix = synthetic.nerf_matrix_to_ngp(synthetic.hemisphere_cameras(7, seed=3)[2])
        tb.fov_axis = 0; tb.fov = math.degrees(synthetic.CAMERA_ANGLE_X)
        tb.render(800, 800, 1, True)
    else:
        for _ in range(n_steps):
            tb.train(); _ = tb.loss
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "timeline.json")
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
k.sort(key=lambda e: e["ts"])
t0 = k[0]["ts"]
for e in k:
    print(f'{e["ts"] - t0:10.1f} {e["dur"]:8.1f} us  stream {e["args"].get("stream")}  {e["name"][:60]}')
