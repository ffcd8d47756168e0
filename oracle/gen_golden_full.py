"""TEST INFRASTRUCTURE ONLY. Drives the UNMODIFIED reference `ngp::Testbed` (oracle/_ref/libref_full.so, compiled from /root/reference by
oracle/Makefile.full) on a B200 and stores what it produces as fixtures:

    gpurun -- 'python oracle/gen_golden_full.py gpurun_out/golden_full [small] [big]'

  small  a reference-trained snapshot of a small scene written by the reference's own save_snapshot (configs/nerf/base.json with
         log2_hashmap_size 15 so that the file stays small), frames rendered from it by the reference's render_frame (classic path,
         K17 + accumulate + tonemap) and bl_render_frame (Blender path, K18: instances, opacity, masks, camera models, depth of field),
         and the occupancy bitfield the reference derives from the snapshot's density grid.
         -> ref_small.msgpack.gz + ref_full_small.npz; copy both to tests/golden/.
  big    BASELINE config 2 at full size (100 x 800^2, batch 2^18): the reference's training speed on this GPU, its PSNR on held-out views,
         and this repo's render of the REFERENCE's snapshot next to the reference's own render of it (PSNR / L1 between the two).
         -> ref_full_big.json (numbers only; copy to profiles/).
"""
import ctypes as C
import gzip
import hashlib
import json
import math
import os
import shutil
import sys
import time
import traceback

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Mask(C.Structure):  # reff_mask
    _fields_ = [("shape", C.c_int), ("mode", C.c_int), ("transform", C.c_float * 16), ("feather", C.c_float), ("opacity", C.c_float), ("dims", C.c_float * 3)]


class Nerf(C.Structure):  # reff_nerf
    _fields_ = [("snapshot_path", C.c_char_p), ("aabb", C.c_float * 6), ("transform", C.c_float * 16), ("opacity", C.c_float), ("n_masks", C.c_int), ("masks", C.POINTER(Mask))]


class Request(C.Structure):  # reff_request
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("mip", C.c_int), ("flip_y", C.c_int), ("spp", C.c_int), ("color_space", C.c_int), ("tonemap_curve", C.c_int),
                ("exposure", C.c_float), ("background_color", C.c_float * 4), ("camera", C.c_float * 12), ("camera_model", C.c_int), ("focal_length", C.c_float),
                ("near_distance", C.c_float), ("aperture_size", C.c_float), ("focus_z", C.c_float), ("spherical_quadrilateral", C.c_float * 3),
                ("quadrilateral_hexahedron", C.c_float * 24), ("aabb", C.c_float * 6), ("n_masks", C.c_int), ("masks", C.POINTER(Mask)), ("n_nerfs", C.c_int), ("nerfs", C.POINTER(Nerf))]


class Ref:
    """ctypes view of oracle/ref_harness/ref_full.cu"""

    def __init__(self):
        self.l = C.CDLL(os.path.join(HERE, "_ref", "libref_full.so"))
        self.l.reff_last_error.restype = C.c_char_p
        self.l.reff_training_step.restype = C.c_uint32
        self.l.reff_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        self.h = C.c_void_p()
        self.ck(self.l.reff_create(C.byref(self.h), 0))  # ETestbedMode::Nerf

    def ck(self, st):
        if st != 0:
            raise RuntimeError("reference: " + self.l.reff_last_error().decode())

    def set(self, **kw):
        for k, v in kw.items():
            self.ck(self.l.reff_set_option(self.h, k.encode(), float(v)))

    def load(self, path):
        self.ck(self.l.reff_load_training_data(self.h, path.encode()))

    def network(self, cfg):
        self.ck(self.l.reff_reload_network_from_json(self.h, json.dumps(cfg).encode()))

    def train(self, batch, n):
        loss = C.c_float(0)
        self.ck(self.l.reff_train(self.h, batch, n, C.byref(loss)))
        return loss.value

    def stats(self):
        s = (C.c_uint64 * 4)()
        self.ck(self.l.reff_stats(self.h, s))
        return dict(rays_per_batch=int(s[0]), measured_batch_size=int(s[1]), measured_batch_size_before_compaction=int(s[2]), n_params=int(s[3]))

    def save(self, path, with_optimizer=False):
        self.ck(self.l.reff_save_snapshot(self.h, path.encode(), int(with_optimizer)))

    def load_snapshot(self, path):
        self.ck(self.l.reff_load_snapshot(self.h, path.encode()))

    def render(self, nerf_c2w, w, h, spp, linear):
        """nerf_c2w: NeRF-convention camera-to-world (set_nerf_camera_matrix converts it with the dataset's scale / offset). Returns (frame, ngp camera 3x4)."""
        out = np.zeros((h, w, 4), np.float32)
        cam = np.ascontiguousarray(np.asarray(nerf_c2w, np.float32)[:3, :4].T.reshape(-1))
        self.ck(self.l.reff_render(self.h, cam.ctypes.data_as(C.c_void_p), w, h, spp, int(linear), out.ctypes.data_as(C.c_void_p)))
        ngp = np.zeros(12, np.float32)
        self.ck(self.l.reff_get_camera(self.h, ngp.ctypes.data_as(C.c_void_p)))
        return out, ngp.reshape(4, 3).T.copy()

    def bl_render(self, rq):
        out = np.zeros((rq.height, rq.width, 4), np.float32)
        self.ck(self.l.reff_bl_render(self.h, C.byref(rq), out.ctypes.data_as(C.c_void_p)))
        return out

    def density_grid(self, n_cascades):
        g = np.zeros(128 ** 3 * n_cascades, np.float32); b = np.zeros(128 ** 3, np.uint8)
        self.ck(self.l.reff_get_density_grid(self.h, g.ctypes.data_as(C.c_void_p), g.size, b.ctypes.data_as(C.c_void_p), b.size))
        return g, b

    def params(self, n):
        w = np.zeros(n, np.float32); e = np.zeros(n, np.float16)
        self.ck(self.l.reff_get_params(self.h, w.ctypes.data_as(C.c_void_p), e.ctypes.data_as(C.c_void_p), n))
        return w, e


def base_config(log2_hashmap_size=19):
    import pyngp
    cfg = json.loads(json.dumps(pyngp.BASE_NETWORK_CONFIG))
    cfg["encoding"]["log2_hashmap_size"] = log2_hashmap_size
    return cfg


def fill(arr, values):
    for i, v in enumerate(values):
        arr[i] = float(v)


def colmajor(m):
    return np.asarray(m, np.float32).T.reshape(-1)


def make_request(w, h, cam_ngp34, focal, nerfs, mip=0, flip_y=0, color_space=0, exposure=0.0, background=(0, 0, 0, 0), near=0.0, aperture=0.0, focus_z=1.0,
                 model=0, sq=(0, 0, 0), qh=None, masks=()):
    rq = Request()
    rq.width, rq.height, rq.mip, rq.flip_y, rq.spp, rq.color_space, rq.tonemap_curve = w, h, mip, flip_y, 1, color_space, 0
    rq.exposure = exposure
    fill(rq.background_color, background)
    fill(rq.camera, colmajor(cam_ngp34))
    rq.camera_model, rq.focal_length, rq.near_distance, rq.aperture_size, rq.focus_z = model, focal, near, aperture, focus_z
    fill(rq.spherical_quadrilateral, sq)
    fill(rq.quadrilateral_hexahedron, np.zeros(24) if qh is None else np.asarray(qh).reshape(-1))
    fill(rq.aabb, (-8, -8, -8, 8, 8, 8))
    rq._keep = [nerfs, masks]
    rq.n_masks = len(masks)
    if masks:
        arr = (Mask * len(masks))(*masks); rq._keep.append(arr); rq.masks = arr
    arr = (Nerf * len(nerfs))(*nerfs); rq._keep.append(arr); rq.nerfs = arr; rq.n_nerfs = len(nerfs)
    return rq


def make_mask(shape, mode, transform, feather, opacity, dims):
    m = Mask()
    m.shape, m.mode, m.feather, m.opacity = shape, mode, feather, opacity
    fill(m.transform, colmajor(transform)); fill(m.dims, list(dims) + [0.0] * (3 - len(dims)))
    return m


def make_nerf(path, transform44=None, opacity=1.0, aabb=(0, 0, 0, 1, 1, 1), masks=()):
    n = Nerf()
    n.snapshot_path = path.encode()
    fill(n.aabb, aabb); fill(n.transform, colmajor(np.eye(4) if transform44 is None else transform44)); n.opacity = opacity
    n.n_masks = len(masks)
    if masks:
        n._keep = (Mask * len(masks))(*masks); n.masks = n._keep
    return n


def translate(x, y, z, s=1.0):
    m = np.eye(4, dtype=np.float32) * s; m[3, 3] = 1.0; m[:3, 3] = (x, y, z)
    return m


SMALL = dict(n_images=24, res=200, steps=1200, batch=1 << 16, log2_hashmap_size=15, frame=128)


def blender_cases(snapshot_path, cam_ngp, focal):
    """name -> kwargs of make_request; shared with the tests through the npz (every request parameter is stored next to the frame)."""
    box_mask = dict(shape=0, mode=0, transform=translate(0.47, 0.40, 0.5), feather=0.0, opacity=1.0, dims=(0.30, 0.22, 0.30))
    sphere_sub = dict(shape=2, mode=1, transform=translate(0.45, 0.40, 0.5), feather=0.06, opacity=1.0, dims=(0.10,))
    cyl_add = dict(shape=1, mode=0, transform=translate(0.5, 0.42, 0.5), feather=0.04, opacity=0.8, dims=(0.16, 0.5))
    qh = np.array([[-0.05, 0.05, 0.1], [0.05, 0.05, 0.1], [-0.05, -0.05, 0.1], [0.05, -0.05, 0.1],     # front tl tr bl br
                   [-0.02, 0.02, 0.0], [0.02, 0.02, 0.0], [-0.02, -0.02, 0.0], [0.02, -0.02, 0.0]], np.float32)  # back
    return {
        "single": dict(nerfs=[dict()]),
        "two_instances": dict(nerfs=[dict(), dict(transform=translate(0.25, 0.05, -0.1, 0.8), opacity=0.6)], background=(0.1, 0.2, 0.3, 1.0)),
        "mip1_flip_srgb": dict(nerfs=[dict()], mip=1, flip_y=1, color_space=1, background=(0.3, 0.3, 0.3, 0.5), exposure=0.5),
        "mask_box_add": dict(nerfs=[dict(masks=[box_mask])]),
        "mask_sphere_subtract_global": dict(nerfs=[dict()], masks=[sphere_sub]),
        "mask_cylinder_feather": dict(nerfs=[dict(masks=[cyl_add])], near=0.05),
        "camera_spherical_quadrilateral": dict(nerfs=[dict()], model=2, sq=(0.35, 0.35, 0.12)),
        "camera_quadrilateral_hexahedron": dict(nerfs=[dict()], model=1, qh=qh),
        "depth_of_field": dict(nerfs=[dict()], aperture=0.02, focus_z=1.2),
    }


def build_request(case, snapshot_path, w, h, cam_ngp, focal):
    kw = dict(case)
    nerfs = [make_nerf(snapshot_path, n.get("transform"), n.get("opacity", 1.0), n.get("aabb", (0, 0, 0, 1, 1, 1)), [make_mask(**m) for m in n.get("masks", [])]) for n in kw.pop("nerfs")]
    masks = [make_mask(**m) for m in kw.pop("masks", [])]
    return make_request(w, h, cam_ngp, focal, nerfs, masks=masks, **kw)


def gen_small(out_dir):
    import synthetic
    scratch = "/tmp/ngpb_ref_small"
    shutil.rmtree(scratch, ignore_errors=True)
    scene = synthetic.make_lego_scene(SMALL["n_images"], SMALL["res"], seed=0)
    tj = synthetic.write_transforms_json(scene, scratch)
    ref = Ref()
    ref.load(tj)
    ref.network(base_config(SMALL["log2_hashmap_size"]))
    ref.set(shall_train=1)
    loss = ref.train(SMALL["batch"], SMALL["steps"])
    st = ref.stats()
    print("reference trained the small scene:", SMALL["steps"], "steps, loss", loss, st)
    snap = os.path.join(out_dir, "ref_small.msgpack")
    ref.save(snap, False)
    with open(snap, "rb") as f:
        raw = f.read()
    with gzip.open(snap + ".gz", "wb", compresslevel=9) as f:
        f.write(raw)
    out = dict(steps=SMALL["steps"], batch=SMALL["batch"], loss=np.float32(loss), n_params=st["n_params"], snapshot_sha256=np.frombuffer(hashlib.sha256(raw).digest(), np.uint8))

    # occupancy: what the reference derives from the snapshot's (fp16) density grid after load_snapshot
    ref2 = Ref()
    ref2.load_snapshot(snap)
    grid, bits = ref2.density_grid(1)
    out["bitfield_packed"] = bits  # npz compression handles the sparsity
    out["density_grid_sum"] = np.float64(grid.astype(np.float64).sum())
    _, ema = ref2.params(st["n_params"])
    out["params_sha256"] = np.frombuffer(hashlib.sha256(ema.tobytes()).digest(), np.uint8)

    # classic renders (render_frame -> render_nerf -> accumulate + tonemap) from the loaded snapshot, held-out cameras
    W = H = SMALL["frame"]
    cams = synthetic.hemisphere_cameras(3, seed=7)
    fov_deg = math.degrees(synthetic.CAMERA_ANGLE_X)
    classic = [
        dict(cam=0, spp=1, linear=1, snap=1, min_t=1e-4, bg=(0, 0, 0, 1), exposure=0.0),
        dict(cam=1, spp=4, linear=0, snap=0, min_t=0.01, bg=(0.2, 0.4, 0.6, 1.0), exposure=0.0),
        dict(cam=2, spp=2, linear=1, snap=1, min_t=0.01, bg=(1.0, 1.0, 1.0, 0.5), exposure=1.0),
    ]
    ngp_cams = []
    for i, c in enumerate(classic):
        ref2.set(fov_axis=0, fov=fov_deg, snap_to_pixel_centers=c["snap"], render_min_transmittance=c["min_t"], exposure=c["exposure"],
                 background_color_r=c["bg"][0], background_color_g=c["bg"][1], background_color_b=c["bg"][2], background_color_a=c["bg"][3], dynamic_res=0)
        frame, ngp = ref2.render(cams[c["cam"]], W, H, c["spp"], c["linear"])
        out[f"classic_{i}"] = frame
        out[f"classic_{i}_cfg"] = np.array([c["cam"], c["spp"], c["linear"], c["snap"], c["min_t"], *c["bg"], c["exposure"]], np.float64)
        ngp_cams.append(ngp)
        print(f"classic {i}: mean rgba {frame.reshape(-1, 4).mean(0)}")
    out["nerf_cams"] = np.stack([np.asarray(c, np.float32) for c in cams]); out["ngp_cams"] = np.stack(ngp_cams); out["fov_deg"] = fov_deg

    # Blender path
    focal = 0.5 * W / math.tan(0.5 * synthetic.CAMERA_ANGLE_X)
    out["bl_focal"] = np.float32(focal)
    cases = blender_cases("SNAPSHOT", ngp_cams[0], focal)
    out["bl_cases_json"] = np.frombuffer(json.dumps({k: json.loads(json.dumps(v, default=lambda a: np.asarray(a).tolist())) for k, v in cases.items()}).encode(), np.uint8)
    for name, case in cases.items():
        try:
            rq = build_request(case, snap, W, H, ngp_cams[0], focal)
            frame = ref2.bl_render(rq)
            out[f"bl_{name}"] = frame
            print(f"blender {name}: mean rgba {frame.reshape(-1, 4).mean(0)}")
        except Exception:
            traceback.print_exc()
    # the other direction: this repo loads the reference's snapshot, writes it back with ITS OWN writer, and the reference loads that file
    try:
        import pyngp
        tb = pyngp.Testbed()
        tb.load_snapshot(snap)
        ours_path = "/tmp/ngpb_ref_small/ours_resaved.msgpack"
        tb.save_snapshot(ours_path)
        ref3 = Ref()
        ref3.load_snapshot(ours_path)
        _, ema3 = ref3.params(st["n_params"])
        c = classic[0]
        ref3.set(fov_axis=0, fov=fov_deg, snap_to_pixel_centers=c["snap"], render_min_transmittance=c["min_t"], exposure=c["exposure"],
                 background_color_r=c["bg"][0], background_color_g=c["bg"][1], background_color_b=c["bg"][2], background_color_a=c["bg"][3], dynamic_res=0)
        # no nerf.dataset block in our file -> the reference's set_nerf_camera_matrix has no scale / offset: hand it the ngp matrix through m_camera instead
        frame3 = ref3.bl_render(build_request(cases["single"], ours_path, W, H, ngp_cams[0], focal))
        out["reference_reads_our_snapshot"] = np.array([float(np.array_equal(ema3, ema)), float(np.abs(frame3 - out["bl_single"]).max())], np.float64)
        print("reference loaded the snapshot this repo wrote: params equal", np.array_equal(ema3, ema), "blender frame max abs diff", float(np.abs(frame3 - out["bl_single"]).max()))
    except Exception:
        traceback.print_exc()
    np.savez_compressed(os.path.join(out_dir, "ref_full_small.npz"), **out)
    print("small fixture written:", os.path.getsize(snap + ".gz"), "bytes of snapshot,", os.path.getsize(os.path.join(out_dir, "ref_full_small.npz")), "bytes of frames")


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 10.0 * math.log10(1.0 / max(mse, 1e-12))


def linear_to_srgb(x):
    return np.where(x < 0.0031308, 12.92 * x, 1.055 * np.maximum(x, 1e-8) ** (1 / 2.4) - 0.055)


def gen_big(out_dir, steps=2000, timed=300):
    """BASELINE config 2 at full size, reference and this repo side by side on the same GPU."""
    import torch
    import synthetic
    import pyngp
    scratch = "/tmp/ngpb_ref_big"
    shutil.rmtree(scratch, ignore_errors=True)
    t0 = time.time()
    scene = synthetic.make_lego_scene(100, 800, seed=0)
    tj = synthetic.write_transforms_json(scene, scratch)
    print(f"dataset written in {time.time() - t0:.1f} s")
    report = dict(config="Lego-shaped synthetic scene, 100 x 800^2 RGBA8, configs/nerf/base.json, batch 2^18, seed 1337", steps=steps)
    B = 1 << 18
    ref = Ref()
    t0 = time.time(); ref.load(tj); report["reference_load_s"] = time.time() - t0
    ref.network(base_config(19))
    ref.set(shall_train=1)
    ref.train(B, steps - timed)
    torch.cuda.synchronize()
    t0 = time.time(); loss = ref.train(B, timed); dt = time.time() - t0  # reff_train ends with cudaDeviceSynchronize
    report["reference"] = dict(it_per_s=timed / dt, ms_per_step=1e3 * dt / timed, loss=loss, timed_steps=timed, after_steps=steps - timed, stats=ref.stats(),
                               how="wall clock around `timed` calls of ngp::Testbed::train(2^18) + cudaDeviceSynchronize, unmodified reference compiled for sm_100")
    print("reference:", report["reference"])
    snap = os.path.join(scratch, "ref_big.msgpack")
    ref.save(snap, False)

    # held-out views (scripts/run.py:216-303): 800^2, 8 spp, snap_to_pixel_centers, min transmittance 1e-4, black background, sRGB PSNR vs ground truth
    cams = synthetic.hemisphere_cameras(4, seed=11)
    fov_deg = math.degrees(synthetic.CAMERA_ANGLE_X)
    fx = 0.5 * 800 / math.tan(0.5 * synthetic.CAMERA_ANGLE_X)
    ref.set(fov_axis=0, fov=fov_deg, snap_to_pixel_centers=1, render_min_transmittance=1e-4, exposure=0.0, background_color_r=0, background_color_g=0, background_color_b=0, background_color_a=1, dynamic_res=0)
    tb = pyngp.Testbed()
    tb.load_snapshot(snap)  # THIS repo renders the REFERENCE's weights
    tb.fov_axis = 0; tb.fov = fov_deg; tb.snap_to_pixel_centers = True; tb.nerf.render_min_transmittance = 1e-4; tb.background_color = [0, 0, 0, 1]
    ours = pyngp.Testbed()
    ours.load_training_data(tj)
    ours.train_n(steps, B)
    ours.fov_axis = 0; ours.fov = fov_deg; ours.snap_to_pixel_centers = True; ours.nerf.render_min_transmittance = 1e-4; ours.background_color = [0, 0, 0, 1]
    rows = []
    for i, c in enumerate(cams):
        gt8 = synthetic.render_image(synthetic.nerf_matrix_to_ngp(c), 800, fx, fx, synthetic.lego_boxes(), device="cuda").cpu().numpy().astype(np.float32) / 255.0
        gt = gt8[..., :3] * gt8[..., 3:4]  # composited on black, sRGB
        fr, ngp = ref.render(c, 800, 800, 8, 0)
        tb.set_nerf_camera_matrix(np.asarray(c)[:3]); fo = tb.render(800, 800, 8, linear=False)
        ours.set_nerf_camera_matrix(np.asarray(c)[:3]); ft = ours.render(800, 800, 8, linear=False)
        clip = lambda f: np.clip(f[..., :3], 0, 1)
        rows.append(dict(view=i, psnr_reference_vs_gt=psnr(clip(fr), gt), psnr_ours_trained_vs_gt=psnr(clip(ft), gt), psnr_ours_on_reference_weights_vs_gt=psnr(clip(fo), gt),
                         psnr_ours_vs_reference_same_weights=psnr(fo, fr), l1_ours_vs_reference_same_weights=float(np.abs(fo - fr).mean()),
                         max_abs_ours_vs_reference_same_weights=float(np.abs(fo - fr).max())))
        print(rows[-1])
    report["held_out_views"] = rows
    for k in rows[0]:
        if k != "view":
            report["mean_" + k] = float(np.mean([r[k] for r in rows]))
    with open(os.path.join(out_dir, "ref_full_big.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps({k: v for k, v in report.items() if k.startswith("mean_")}))


def _pose_errors(get, truth):
    pos, ang = [], []
    for i, t in enumerate(truth):
        m = np.asarray(get(i), np.float64)
        pos.append(np.linalg.norm(m[:, 3] - t[:, 3]))
        r = m[:, :3] @ np.asarray(t, np.float64)[:, :3].T
        ang.append(np.degrees(np.arccos(np.clip((np.trace(r) - 1) / 2, -1, 1))))
    return dict(mean_position_error=float(np.mean(pos)), mean_rotation_error_deg=float(np.mean(ang)))


def gen_config3(out_dir, steps=3000):
    """BASELINE config 3 (occupancy-grid update + camera-extrinsics optimisation), reference and this repo side by side on the same GPU.
      A. the synthetic scene with every training camera perturbed (1.7 degrees, 0.03 units): pose error against the true cameras before / after training with
         nerf.training.optimize_extrinsics, for both implementations (same protocol, same dataset files);
      B. the reference's bundled fox capture (50 images 1080 x 1920, OpenCV lens, aabb_scale 4), when a copy lies under oracle/_ref/fox (git-ignored; it
         travels to the GPU box): training speed and loss with the option on, both implementations.
    -> ref_config3.json (numbers only; copy to profiles/)."""
    import torch
    import synthetic
    import pyngp
    report = dict(steps=steps)
    # ---- A ----
    scratch = "/tmp/ngpb_ref_c3"
    shutil.rmtree(scratch, ignore_errors=True)
    n_cam, res, B = 50, 400, 1 << 16
    scene = synthetic.make_lego_scene(n_cam, res, seed=0)
    truth = [np.asarray(c, np.float32)[:3] for c in scene["nerf_c2w"]]
    rs = np.random.RandomState(11)
    perturbed = []
    for t in truth:
        axis = rs.randn(3); axis /= np.linalg.norm(axis)
        ang = np.radians(1.7)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        p = np.eye(4, dtype=np.float32)
        p[:3, :3] = (R @ t[:, :3]).astype(np.float32)
        p[:3, 3] = t[:, 3] + (rs.randn(3) * 0.03).astype(np.float32)
        perturbed.append(p)
    scene_p = dict(scene); scene_p["nerf_c2w"] = [p.tolist() for p in perturbed]
    tj = synthetic.write_transforms_json(scene_p, scratch)
    # both loaders sort the frames by file_path (nerf_loader.cu:356-358): frame index i of the testbed = i-th name in lexicographic order
    order = sorted(range(n_cam), key=lambda i: f"./train/r_{i}")
    truth = [truth[i] for i in order]
    A = dict(config=f"Lego-shaped synthetic scene, {n_cam} x {res}^2, cameras perturbed by 1.7 deg / sigma 0.03, batch 2^16, {steps} steps, optimize_extrinsics")
    ref = Ref()
    ref.l.reff_get_camera_extrinsics.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    def ref_get(i):
        o = np.zeros(12, np.float32)
        ref.ck(ref.l.reff_get_camera_extrinsics(ref.h, i, o.ctypes.data))
        return o.reshape(4, 3).T
    ref.load(tj); ref.network(base_config(19)); ref.set(shall_train=1, optimize_extrinsics=1)
    A["before"] = _pose_errors(ref_get, truth)
    t0 = time.time(); loss = ref.train(B, steps); dt = time.time() - t0
    A["reference"] = dict(_pose_errors(ref_get, truth), loss=loss, it_per_s=steps / dt)
    ours = pyngp.Testbed()
    ours.load_training_data(tj)
    ours.nerf.training.optimize_extrinsics = True
    A["before_ours"] = _pose_errors(ours.nerf.training.get_camera_extrinsics, truth)
    torch.cuda.synchronize(); t0 = time.time(); ours.train_n(steps, B); dt = time.time() - t0
    A["ours"] = dict(_pose_errors(ours.nerf.training.get_camera_extrinsics, truth), loss=ours.loss, it_per_s=steps / dt)
    print("config 3 / A:", json.dumps(A))
    report["perturbed_synthetic"] = A
    del ref, ours
    # ---- B ----
    fox = os.path.join(HERE, "_ref", "fox", "transforms.json")
    if os.path.exists(fox):
        Bf, n_steps, timed = 1 << 18, 2000, 300
        Bd = dict(config="data/nerf/fox of the reference (50 x 1080 x 1920 JPEG, OpenCV lens, aabb_scale 4), configs/nerf/base.json, batch 2^18, optimize_extrinsics, 2000 steps")
        ref = Ref()
        ref.l.reff_get_camera_extrinsics.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        ref.load(fox); ref.network(base_config(19)); ref.set(shall_train=1, optimize_extrinsics=1)
        start = []
        for i in range(50):
            o = np.zeros(12, np.float32); ref.ck(ref.l.reff_get_camera_extrinsics(ref.h, i, o.ctypes.data)); start.append(o.reshape(4, 3).T.copy())
        ref.train(Bf, n_steps - timed)
        t0 = time.time(); loss = ref.train(Bf, timed); dt = time.time() - t0
        moved = []
        for i in range(50):
            o = np.zeros(12, np.float32); ref.ck(ref.l.reff_get_camera_extrinsics(ref.h, i, o.ctypes.data)); moved.append(o.reshape(4, 3).T.copy())
        Bd["reference"] = dict(it_per_s=timed / dt, ms_per_step=1e3 * dt / timed, loss=loss, stats=ref.stats(), camera_motion=_pose_errors(lambda i: moved[i], start))
        del ref
        ours = pyngp.Testbed()
        ours.load_training_data(fox)
        ours.nerf.training.optimize_extrinsics = True
        start = [ours.nerf.training.get_camera_extrinsics(i) for i in range(50)]
        ours.train_n(n_steps - timed, Bf)
        torch.cuda.synchronize(); t0 = time.time(); ours.train_n(timed, Bf); dt = time.time() - t0
        Bd["ours"] = dict(it_per_s=timed / dt, ms_per_step=1e3 * dt / timed, loss=ours.loss, stats=ours.stats(),
                          camera_motion=_pose_errors(ours.nerf.training.get_camera_extrinsics, start))
        print("config 3 / B:", json.dumps(Bd))
        report["fox"] = Bd
    else:
        report["fox"] = "oracle/_ref/fox absent"
    with open(os.path.join(out_dir, "ref_config3.json"), "w") as f:
        json.dump(report, f, indent=1)


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def gen_modes(out_dir):
    """ETestbedMode::Image and ::Sdf of the unmodified reference on the inputs of tests/golden_inputs.py (procedural_image, sdf_pool):
    initial parameters, the training batches of the first steps, the loss every 16th step, and for the image model the trained parameters with the
    reference's own inference, compute_image_mse and render of them. -> ref_modes.npz (copy to tests/golden/)."""
    from PIL import Image as PILImage
    from golden_inputs import procedural_image, sdf_pool, MODE_IMAGE_RES, MODE_IMAGE_BATCH, MODE_SDF_BATCH
    scratch = "/tmp/ngpb_ref_modes"
    shutil.rmtree(scratch, ignore_errors=True); os.makedirs(scratch)
    out = {}

    def mode_ref(mode):
        r = Ref.__new__(Ref)
        r.l = C.CDLL(os.path.join(HERE, "_ref", "libref_full.so"))
        r.l.reff_last_error.restype = C.c_char_p
        r.l.reff_training_step.restype = C.c_uint32
        r.l.reff_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        r.h = C.c_void_p()
        r.ck(r.l.reff_create(C.byref(r.h), mode))
        return r

    def params(r):
        n = r.stats()["n_params"]
        fp = np.zeros(n, np.float32); hf = np.zeros(n, np.float16)
        r.ck(r.l.reff_get_params(r.h, fp.ctypes.data_as(C.c_void_p), hf.ctypes.data_as(C.c_void_p), n))
        return fp, hf

    # ---- image ----
    w, h = MODE_IMAGE_RES
    png = os.path.join(scratch, "image.png")
    PILImage.fromarray(procedural_image()).save(png)
    from pyngp.modes import IMAGE_NETWORK_CONFIG, SDF_NETWORK_CONFIG
    r = mode_ref(2)
    r.load(png); r.network(IMAGE_NETWORK_CONFIG); r.set(shall_train=1)
    wh = (C.c_int * 2)(); r.ck(r.l.reff_image_resolution(r.h, wh)); assert (wh[0], wh[1]) == (w, h)
    data = np.zeros((h, w, 4), np.float32); r.ck(r.l.reff_image_data(r.h, data.ctypes.data_as(C.c_void_p)))
    out["image_data_sha"] = _sha(data); out["image_data_head"] = data.reshape(-1, 4)[:4096].copy()
    fp, _ = params(r)
    out["image_n_params"] = np.uint32(fp.shape[0]); out["image_init_sha"] = _sha(fp); out["image_init_head"] = fp[:8192].copy(); out["image_init_grid_head"] = fp[7168:7168 + 4096].copy()
    B = MODE_IMAGE_BATCH
    losses = []
    for step in range(3):
        losses.append(r.train(B, 1))
        pos = np.zeros((B, 2), np.float32); tgt = np.zeros((B, 3), np.float32)
        r.ck(r.l.reff_image_training_batch(r.h, B, pos.ctypes.data_as(C.c_void_p), tgt.ctypes.data_as(C.c_void_p)))
        out[f"image_batch{step}_pos_sha"] = _sha(pos); out[f"image_batch{step}_tgt_sha"] = _sha(tgt)
        out[f"image_batch{step}_pos_head"] = pos[:2048].copy(); out[f"image_batch{step}_tgt_head"] = tgt[:2048].copy()
    # reff_train returns m_loss_scalar.val(): the loss of the last step whose index was a multiple of 16 (the reference reads the scalar on those steps only)
    curve = [losses[0]]
    for k in range(1, 63):
        curve.append(r.train(B, 16 if k > 1 else 14))  # up to and including step 16 k
    out["image_loss_curve"] = np.array(curve, np.float32); out["image_steps"] = np.uint32(r.l.reff_training_step(r.h))
    _, hf = params(r)
    out["image_trained_params"] = hf
    mse = C.c_float()
    r.ck(r.l.reff_image_mse(r.h, 0, C.byref(mse))); out["image_mse"] = np.float32(mse.value)
    r.ck(r.l.reff_image_mse(r.h, 1, C.byref(mse))); out["image_mse_quantized"] = np.float32(mse.value)
    ys, xs = np.meshgrid(np.arange(64), np.arange(64), indexing="ij")
    q = np.stack([(xs.reshape(-1) * 8 + 0.5) / w, (ys.reshape(-1) * 6 + 0.5) / h], 1).astype(np.float32)
    o = np.zeros((4096, 3), np.float32)
    r.ck(r.l.reff_inference(r.h, q.ctypes.data_as(C.c_void_p), 4096, 2, 3, o.ctypes.data_as(C.c_void_p)))
    out["image_query"] = q; out["image_inference"] = o
    r.set(snap_to_pixel_centers=1, dynamic_res=0, background_color_r=0.2, background_color_g=0.3, background_color_b=0.4, background_color_a=1.0, exposure=0.0)
    for name, (rw, rh, spp, linear) in dict(a=(256, 192, 1, 0), b=(200, 120, 2, 1)).items():
        fr = np.zeros((rh, rw, 4), np.float32)
        r.ck(r.l.reff_render(r.h, None, rw, rh, spp, linear, fr.ctypes.data_as(C.c_void_p)))
        out[f"image_render_{name}"] = fr.astype(np.float16); out[f"image_render_{name}_cfg"] = np.array([rw, rh, spp, linear], np.int32)
    print("image: loss", curve[0], "->", curve[-1], "mse", out["image_mse"], "steps", out["image_steps"])
    del r

    # ---- sdf ----
    bunny = os.path.join(HERE, "_ref", "bunny.obj")
    r = mode_ref(1)
    r.load(bunny); r.network(SDF_NETWORK_CONFIG); r.set(shall_train=1)
    info = np.zeros(7, np.float32); r.ck(r.l.reff_sdf_mesh_info(r.h, info.ctypes.data_as(C.c_void_p))); out["sdf_bunny_mesh_info"] = info
    fp, _ = params(r)
    out["sdf_n_params"] = np.uint32(fp.shape[0]); out["sdf_init_sha"] = _sha(fp); out["sdf_init_head"] = fp[:8192].copy()
    pos, dist = sdf_pool()
    r.ck(r.l.reff_sdf_override_training_data(r.h, pos.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), pos.shape[0]))
    B = MODE_SDF_BATCH
    curve = []
    for step in range(3):
        curve.append(r.train(B, 1))
        bp = np.zeros((B, 3), np.float32); bd = np.zeros(B, np.float32)
        r.ck(r.l.reff_sdf_training_batch(r.h, B, bp.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p)))
        out[f"sdf_batch{step}_pos_sha"] = _sha(bp); out[f"sdf_batch{step}_dist_sha"] = _sha(bd); out[f"sdf_batch{step}_dist_head"] = bd[:1024].copy()
    for k in range(1, 63):
        curve.append(r.train(B, 16 if k > 1 else 14))
    out["sdf_loss_curve"] = np.array(curve[0:1] + curve[3:], np.float32); out["sdf_steps"] = np.uint32(r.l.reff_training_step(r.h))
    print("sdf: loss", curve[0], "->", curve[-1], "steps", out["sdf_steps"], "bunny", info)
    np.savez_compressed(os.path.join(out_dir, "ref_modes.npz"), **out)


def gen_modes_speed(out_dir, steps=600, timed=400):
    """BASELINE configs 1 and 5 as training throughput, reference and this repo side by side on the same GPU: neural image 512 x 512 (procedural), SDF on a
    supplied pool of 2^20 analytic pairs; batch 2^18 (the Testbed default), wall clock around `timed` steps after a warm-up. -> ref_modes_speed.json"""
    import torch
    import pyngp
    from PIL import Image as PILImage
    from pyngp.modes import IMAGE_NETWORK_CONFIG, SDF_NETWORK_CONFIG
    import golden_inputs as gi
    scratch = "/tmp/ngpb_ref_modes_speed"
    shutil.rmtree(scratch, ignore_errors=True); os.makedirs(scratch)
    B = 1 << 18
    report = dict(batch=B, steps=steps, timed=timed, how="wall clock around `timed` calls of Testbed::train(2^18) ending in a device synchronisation")
    gi.MODE_IMAGE_RES = (512, 512)
    img = gi.procedural_image()
    png = os.path.join(scratch, "image.png"); PILImage.fromarray(img).save(png)
    rs = np.random.RandomState(1)
    pool = rs.rand(1 << 20, 3).astype(np.float32)
    dist = (np.linalg.norm(pool - 0.5, axis=1) - 0.3).astype(np.float32)

    def ref_of(mode):
        r = Ref.__new__(Ref)
        r.l = C.CDLL(os.path.join(HERE, "_ref", "libref_full.so")); r.l.reff_last_error.restype = C.c_char_p
        r.l.reff_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        r.h = C.c_void_p(); r.ck(r.l.reff_create(C.byref(r.h), mode))
        return r
    for name, mode in (("image", 2), ("sdf", 1)):
        r = ref_of(mode)
        if name == "image":
            r.load(png); r.network(IMAGE_NETWORK_CONFIG)
        else:
            r.load(os.path.join(HERE, "_ref", "bunny.obj")); r.network(SDF_NETWORK_CONFIG)
            r.ck(r.l.reff_sdf_override_training_data(r.h, pool.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), pool.shape[0]))
        r.set(shall_train=1)
        r.train(B, steps - timed)
        t0 = time.time(); loss = r.train(B, timed); dt = time.time() - t0
        ref_row = dict(it_per_s=timed / dt, ms_per_step=1e3 * dt / timed, loss=loss)
        del r
        tb = pyngp.Testbed(pyngp.TestbedMode.Image if name == "image" else pyngp.TestbedMode.Sdf)
        if name == "image":
            tb.load_image_data(img)
        else:
            tb.set_unit_cube_pairs(pool, dist)
        tb.train_n(steps - timed, B)
        torch.cuda.synchronize(); tb.training_batch(0)
        t0 = time.time(); tb.train_n(timed, B); tb.training_batch(0); dt = time.time() - t0
        ours = dict(it_per_s=timed / dt, ms_per_step=1e3 * dt / timed, loss=tb.loss)
        report[name] = dict(reference=ref_row, ours=ours, speedup=ours["it_per_s"] / ref_row["it_per_s"])
        print(name, json.dumps(report[name]))
    with open(os.path.join(out_dir, "ref_modes_speed.json"), "w") as f:
        json.dump(report, f, indent=1)


def gen_exposure(out_dir):
    """Per-image exposure optimisation (nerf.training.optimize_exposure) of the REFERENCE on the small synthetic scene whose images carry the exposure
    errors of tests/golden_inputs.py:exposure_scene_offsets -> ref_exposure_train.npz: the learned exposures every 500 steps (frame order = the loaders'
    file_path order) and the loss. This repo's run of the same protocol is printed next to it."""
    import synthetic
    import pyngp
    from golden_inputs import EXPOSURE_SCENE, exposure_scene_offsets, apply_image_exposures
    n, res, B, steps = EXPOSURE_SCENE["n_images"], EXPOSURE_SCENE["res"], EXPOSURE_SCENE["batch"], EXPOSURE_SCENE["steps"]
    scratch = "/tmp/ngpb_ref_exposure"
    shutil.rmtree(scratch, ignore_errors=True)
    scene = synthetic.make_lego_scene(n, res, seed=0)
    e = exposure_scene_offsets(n)
    scene = dict(scene); scene["images"] = apply_image_exposures(np.asarray(scene["images"]), e)
    tj = synthetic.write_transforms_json(scene, scratch)
    order = sorted(range(n), key=lambda i: f"./train/r_{i}")
    ref = Ref()
    ref.load(tj)
    ref.network(base_config())
    ref.set(shall_train=1, optimize_exposure=1)
    learned, losses = [], []
    for _ in range(steps // 500):
        losses.append(ref.train(B, 500))
        o = np.zeros((n, 3), np.float32); ref.ck(ref.l.reff_get_exposures(ref.h, o.ctypes.data_as(C.c_void_p))); learned.append(o)
    want = -(e[order] - e[order].mean(0))
    print("reference: loss", losses, "rms(learned - expected)", float(np.sqrt(np.mean((learned[-1] - want) ** 2))), "rms(expected)", float(np.sqrt(np.mean(want ** 2))))
    np.savez_compressed(os.path.join(out_dir, "ref_exposure_train.npz"), offsets=e, order=np.array(order, np.int32), learned=np.stack(learned), losses=np.array(losses, np.float32))
    tb = pyngp.Testbed()
    tb.load_training_data(tj)
    tb.nerf.training.optimize_exposure = True
    for k in range(steps // 500):
        tb.train_n(500, B)
        o = tb.nerf.training.get_camera_exposures()
        print(f"ours after {tb.training_step}: loss {tb.loss:.6f} rms(ours - reference) {float(np.sqrt(np.mean((o - learned[k]) ** 2))):.4f} rms(ours - expected) {float(np.sqrt(np.mean((o - want) ** 2))):.4f}"
              f" corr {float(np.corrcoef(o.ravel(), learned[k].ravel())[0, 1]):.4f}")


def gen_error_map(out_dir):
    """K19 through the REFERENCE's own Testbed: the small scene with one damaged image (tests/golden_inputs.py:error_scene_images), trained with
    sample_focal_plane_proportional_to_error + sample_image_proportional_to_error over three error-map windows (128, 192, 288 steps) ->
    ref_error_map_train.npz: after each window the error-map resolution, the window length, is_cdf_valid, the image probabilities (pmf_img_cpu) and the
    loss. This repo's run of the same protocol is printed next to it."""
    import synthetic
    import pyngp
    from golden_inputs import ERROR_SCENE, error_scene_images
    n, res, B = ERROR_SCENE["n_images"], ERROR_SCENE["res"], ERROR_SCENE["batch"]
    scratch = "/tmp/ngpb_ref_error_map"
    shutil.rmtree(scratch, ignore_errors=True)
    scene = dict(synthetic.make_lego_scene(n, res, seed=0))
    scene["images"] = error_scene_images(np.asarray(scene["images"]))
    tj = synthetic.write_transforms_json(scene, scratch)
    ref = Ref()
    ref.load(tj)
    ref.network(base_config())
    ref.set(shall_train=1, sample_focal_plane_proportional_to_error=1, sample_image_proportional_to_error=1)
    states, pmfs, losses, rpb = [], [], [], []
    for w in ERROR_SCENE["windows"]:
        losses.append(ref.train(B, w))
        st = (C.c_int * 5)(); pmf = np.zeros(n, np.float32)
        ref.ck(ref.l.reff_error_map_state(ref.h, st, pmf.ctypes.data_as(C.c_void_p)))
        states.append(list(st)); pmfs.append(pmf); rpb.append(ref.stats()["rays_per_batch"])
        print("reference after", sum(ERROR_SCENE["windows"][:len(states)]), "steps: state", list(st), "pmf", np.round(pmf, 4), "loss", losses[-1], "rays/batch", rpb[-1])
    np.savez_compressed(os.path.join(out_dir, "ref_error_map_train.npz"), states=np.array(states, np.int32), pmf=np.stack(pmfs), losses=np.array(losses, np.float32),
                        rays_per_batch=np.array(rpb, np.int32))
    tb = pyngp.Testbed()
    tb.load_training_data(tj)
    tr = tb.nerf.training
    tr.sample_focal_plane_proportional_to_error = True
    tr.sample_image_proportional_to_error = True
    for k, w in enumerate(ERROR_SCENE["windows"]):
        tb.train_n(w, B)
        pmf = tr.get_error_map_pmf()
        print(f"ours after {tb.training_step}: res {int(tb._get('error_map_res'))} between {tr.n_steps_between_error_map_updates} since {tr.n_steps_since_error_map_update} pmf {np.round(pmf, 4)}"
              f" loss {tb.loss:.6f} max |pmf - reference| {float(np.abs(pmf - pmfs[k]).max()):.4f}")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden_full")
    os.makedirs(out, exist_ok=True)
    jobs = sys.argv[2:] or ["small", "big"]
    for j in jobs:
        try:
            dict(small=gen_small, big=gen_big, config3=gen_config3, modes=gen_modes, modes_speed=gen_modes_speed, exposure=gen_exposure, error_map=gen_error_map)[j](out)
        except Exception:
            traceback.print_exc()
