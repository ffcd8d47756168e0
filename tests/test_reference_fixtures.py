"""Parity against the UNMODIFIED reference itself (ngp::Testbed compiled headless from /root/reference by oracle/Makefile.full and run on a B200 by
oracle/gen_golden_full.py). The fixtures under tests/golden/ are what the reference produced:

  ref_small.msgpack.gz   a snapshot written by the reference's own save_snapshot after it trained the small synthetic scene for 1200 steps
                         (configs/nerf/base.json with log2_hashmap_size 15)
  ref_full_small.npz     frames the reference rendered from that snapshot: classic path (Testbed::render_frame: K17 + accumulate + tonemap) and Blender
                         path (Testbed::bl_render_frame: K18 with instances, opacity, masks, the fork's camera models, depth of field), the occupancy
                         bitfield it derived from the snapshot's density grid, and the request parameters of every frame
  ref_density_grid.npz   outputs of the reference's six occupancy-grid kernels (K16) on seeded inputs

CPU tests pin the ORACLE on them (so that the oracle used elsewhere as the checker is itself checked against the reference); GPU tests pin the product.
Tolerances: the product accumulates the MLP in fp32 where the reference accumulates in fp16, and both use fast-math exponentials, so frames agree to
PSNR >= 70 dB for the CUDA path (measured 84-100 dB), >= 45 dB for the CPU oracle (libm instead of fast-math exponentials) rather than bit for bit; integer work (occupancy bits, sample indices, mask culling) is exact."""
import gzip
import hashlib
import json
import math
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fx():
    return np.load(os.path.join(GOLDEN, "ref_full_small.npz"))


@pytest.fixture(scope="module")
def snapshot_path(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("ref") / "ref_small.msgpack")
    with gzip.open(os.path.join(GOLDEN, "ref_small.msgpack.gz"), "rb") as f, open(path, "wb") as o:
        o.write(f.read())
    return path


@pytest.fixture(scope="module")
def snapshot(snapshot_path):
    import msgpack
    import pyngp
    with open(snapshot_path, "rb") as f:
        raw = f.read()
    return raw, pyngp.parse_snapshot(msgpack.unpackb(raw, raw=False, strict_map_key=False))


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 10.0 * math.log10(1.0 / max(mse, 1e-12))


def assert_frames_agree(got, want, min_psnr, what):
    assert got.shape == want.shape
    d = np.abs(got - want)
    p = psnr(got, want)
    assert p >= min_psnr, f"{what}: PSNR {p:.1f} dB < {min_psnr} (mean |diff| {d.mean():.2e}, max {d.max():.3f})"
    assert d.mean() <= 1e-3, f"{what}: mean |diff| {d.mean():.2e}"
    return p


# ------------------------------------------------------------------------------------------------------
# CPU: the snapshot container and the oracle against the reference's outputs
# ------------------------------------------------------------------------------------------------------
def test_reference_written_snapshot_parses(fx, snapshot):
    """A .msgpack written by the reference's Testbed::save_snapshot (src/testbed.cu:3008-3042) is read by pyngp.parse_snapshot: parameter blob (fp16, the
    reference's flat order), density grid, aabb_scale, counters, and the NerfDataset block's scale / offset that set_nerf_camera_matrix needs."""
    raw, snap = snapshot
    assert np.array_equal(np.frombuffer(hashlib.sha256(raw).digest(), np.uint8), fx["snapshot_sha256"])
    assert snap["params_half"].shape[0] == int(fx["n_params"]) == 10240 + 2 * (4096 + 12168 + 29792 + 13 * 32768)
    assert np.array_equal(np.frombuffer(hashlib.sha256(snap["params_half"].tobytes()).digest(), np.uint8), fx["params_sha256"])
    assert snap["aabb_scale"] == 1 and snap["training_step"] == int(fx["steps"]) and snap["density_grid"].shape[0] == 128 ** 3
    assert snap["network_config"]["encoding"]["log2_hashmap_size"] == 15
    assert snap["dataset_transform"] == (pytest.approx(0.33), (0.5, 0.5, 0.5))
    assert snap["optimizer"] is None and abs(snap["loss"] - float(fx["loss"])) < 1e-9
    assert int(snap["rgb"]["rays_per_batch"]) > 4096


def test_oracle_occupancy_bits_match_reference(orc, fx, snapshot):
    """update_density_grid_mean_and_bitfield (src/testbed_nerf.cu:2844-2859, incl. the tcnn::reduce_sum mean): the oracle's mean -> threshold -> bitfield ->
    seven max-pooled mips on the snapshot's density grid equal, bit for bit, what the reference computed after load_snapshot."""
    grid = snap_grid = snapshot[1]["density_grid"]
    bits = orc.bitfield(1, np.ascontiguousarray(snap_grid), orc.density_grid_mean(grid))
    assert np.array_equal(bits, fx["bitfield_packed"])
    assert abs(float(grid.astype(np.float64).sum()) - float(fx["density_grid_sum"])) <= 1e-6 * abs(float(fx["density_grid_sum"]))


def _oracle_model(orc, snap):
    m = orc.model(log2_hashmap_size=15)
    assert m.n_grid_params + 10240 == snap["params_half"].shape[0]
    return m


@pytest.mark.parametrize("i", [0, 1, 2])
def test_oracle_classic_render_matches_reference(orc, fx, snapshot, i):
    """Classic render restatement (orc_render_nerf: NerfTracer + accumulate + tonemap) against Testbed::render_frame on the same snapshot: spp 1 / 4 / 2,
    linear and sRGB output, pixel-centre and jittered sampling, two early-termination thresholds, opaque / translucent backgrounds, exposure."""
    snap = snapshot[1]
    cam_i, spp, linear, snap_px, min_t, r, g, b, a, exposure = fx[f"classic_{i}_cfg"]
    W = H = fx[f"classic_{i}"].shape[0]
    f = 0.5 * W / math.tan(0.5 * math.radians(float(fx["fov_deg"])))
    bits = np.ascontiguousarray(fx["bitfield_packed"])
    cfg = orc.render_config(W, H, f, f, fx["ngp_cams"][i], spp=int(spp), snap=bool(snap_px), min_transmittance=float(min_t), output_srgb=not bool(linear),
                            exposure=float(exposure), background=(r, g, b, a))
    got, n = orc.render_nerf(_oracle_model(orc, snap), snap["params_half"], bits, cfg)
    assert n > 10000
    assert_frames_agree(got, fx[f"classic_{i}"], 45.0, f"oracle classic {i}")


def _blender_cases(fx):
    return json.loads(bytes(fx["bl_cases_json"]).decode())


BL_CASES = ["single", "two_instances", "mip1_flip_srgb", "mask_box_add", "mask_sphere_subtract_global", "mask_cylinder_feather",
            "camera_spherical_quadrilateral", "camera_quadrilateral_hexahedron", "depth_of_field"]


@pytest.mark.parametrize("name", BL_CASES)
def test_oracle_blender_render_matches_reference(orc, fx, snapshot, name):
    """Blender-path restatement (orc_blender_render) against Testbed::bl_render_frame on the same snapshot, incl. what round 1 left out: Mask3D (box add,
    global sphere subtract with feather, cylinder with feather + opacity + near distance), SphericalQuadrilateral / QuadrilateralHexahedron cameras, DoF."""
    snap = snapshot[1]
    case = _blender_cases(fx)[name]
    want = fx[f"bl_{name}"]
    H, W = want.shape[:2]
    m = _oracle_model(orc, snap)
    bits = np.ascontiguousarray(fx["bitfield_packed"])
    nerfs = [dict(model=m, params_half=snap["params_half"], bitfield=bits, aabb_scale=1, transform=np.asarray(n.get("transform", np.eye(4)), np.float32),
                  opacity=n.get("opacity", 1.0), masks=n.get("masks", [])) for n in case["nerfs"]]
    got, n_samples = orc.blender_render(W, H, fx["ngp_cams"][0], float(fx["bl_focal"]), nerfs, mip=case.get("mip", 0), flip_y=bool(case.get("flip_y", 0)),
                                        near_distance=case.get("near", 0.0), color_space=case.get("color_space", 0), exposure=case.get("exposure", 0.0),
                                        background=case.get("background", (0, 0, 0, 0)), camera_model=case.get("model", 0), aperture_size=case.get("aperture", 0.0),
                                        focus_z=case.get("focus_z", 1.0), spherical_quadrilateral=case.get("sq", (0, 0, 0)), quadrilateral_hexahedron=case.get("qh"),
                                        masks=case.get("masks", []))
    assert n_samples > 1000
    assert_frames_agree(got, want, 45.0, f"oracle blender {name}")


def test_masks_change_the_image_as_the_reference_says(fx):
    """Sanity of the fixture itself: an Add box mask keeps less than the unmasked NeRF, a Subtract sphere removes part of it, and the reference read back the
    snapshot THIS repo wrote (save_snapshot -> reference load_snapshot + bl_render_frame) with identical parameters and a bit-identical frame."""
    full = fx["bl_single"][..., 3].sum()
    assert fx["bl_mask_box_add"][..., 3].sum() < 0.8 * full and fx["bl_mask_sphere_subtract_global"][..., 3].sum() < 0.8 * full
    assert fx["bl_mask_box_add"][..., 3].sum() > 0.05 * full
    params_equal, frame_diff = fx["reference_reads_our_snapshot"]
    assert params_equal == 1.0 and frame_diff == 0.0


def test_oracle_density_grid_kernels_match_reference(orc):
    """K16 (src/testbed_nerf.cu:369-610): the oracle against the reference's own kernels run on seeded inputs: untrained-cell marking, cell sampling
    (positions and indices, bit-exact), splat + decayed max, thresholding into the bitfield and its mips (SHA-256 of all 2 MiB)."""
    from golden_inputs import DG_CASCADES, DG_CELLS, DG_SAMPLES, DG_STEP, DG_AABB, density_grid_inputs, density_grid_cameras
    g = np.load(os.path.join(GOLDEN, "ref_density_grid.npz"))
    scene = density_grid_cameras()
    imgs = orc.make_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
    grid = np.zeros(DG_CELLS, np.float32)
    orc.mark_untrained(grid, imgs, clear_visible=True)
    assert np.array_equal(np.packbits(grid < 0), g["untrained_bits"]) and set(np.unique(grid).tolist()) <= {-1.0, 0.0}
    grid_in, density = density_grid_inputs()
    rng = orc.pcg32(4242)
    assert int(g["rng_state"]) == rng.state and int(g["rng_inc"]) == rng.inc
    for name, thresh in (("uniform", -0.01), ("occupied", 0.01)):
        pos, idx = orc.generate_grid_samples(DG_SAMPLES, rng, DG_STEP, DG_AABB, grid_in, DG_CASCADES, thresh)
        assert np.array_equal(idx, g[f"idx_{name}"]), name
        assert np.array_equal(pos.view(np.uint32), g[f"pos_{name}"].view(np.uint32)), name
    ema = grid_in.copy()
    orc.splat_and_ema(g["idx_occupied"], density, 0.95, ema)
    touched = g["ema_touched_idx"]
    np.testing.assert_allclose(ema[touched], g["ema_touched_val"], rtol=2e-5, atol=1e-9)  # __expf in the reference's splat kernel vs libm
    assert int((ema < 0).sum()) == int(g["ema_negative_count"])
    # untouched cells are exact (decay of a float); with the reference's values at the touched cells the grid is the reference's grid
    ema[touched] = g["ema_touched_val"]
    assert np.array_equal(ema[:65536].view(np.uint32), g["ema_head"].view(np.uint32))
    assert abs(float(ema.astype(np.float64).sum()) - float(g["ema_sum"])) <= 1e-9 * abs(float(g["ema_sum"]))
    bits = orc.bitfield(DG_CASCADES, ema, float(g["mean"]))
    assert np.array_equal(bits[:16384], g["bitfield_head"])
    assert [int(np.unpackbits(bits[m * 128 ** 3 // 8:(m + 1) * 128 ** 3 // 8]).sum()) for m in range(8)] == g["bitfield_popcount_per_mip"].tolist()
    assert np.array_equal(np.frombuffer(hashlib.sha256(bits.tobytes()).digest(), np.uint8), g["bitfield_sha256"])


# ------------------------------------------------------------------------------------------------------
# GPU: the product against the reference's outputs
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def loaded_testbed(snapshot_path):
    import pyngp
    tb = pyngp.Testbed()
    tb.load_snapshot(snapshot_path)
    return tb


@pytest.mark.gpu
def test_product_loads_reference_snapshot(fx, loaded_testbed, snapshot):
    """pyngp.Testbed.load_snapshot on the reference-written file: parameters arrive bit for bit, and the occupancy bitfield the CUDA path derives from the
    snapshot's density grid (mean, threshold, 7 mips) equals the reference's bitfield exactly."""
    tb = loaded_testbed
    assert tb.n_params == int(fx["n_params"]) and tb.training_step == int(fx["steps"])
    _, w_half, w_ema = tb.get_params()
    assert np.array_equal(w_ema.view(np.uint16), snapshot[1]["params_half"].view(np.uint16)) and np.array_equal(w_half.view(np.uint16), w_ema.view(np.uint16))
    _, bits = tb.get_density_grid()
    assert np.array_equal(bits, fx["bitfield_packed"])


@pytest.mark.gpu
@pytest.mark.parametrize("i", [0, 1, 2])
def test_product_classic_render_matches_reference(fx, loaded_testbed, i):
    """pyngp.Testbed.render (K17 + accumulate + tonemap) against Testbed::render_frame of the reference on the same snapshot, through the public surface:
    set_nerf_camera_matrix with the dataset transform restored from the snapshot, fov, snap_to_pixel_centers, render_min_transmittance, background, exposure."""
    tb = loaded_testbed
    cam_i, spp, linear, snap_px, min_t, r, g, b, a, exposure = fx[f"classic_{i}_cfg"]
    want = fx[f"classic_{i}"]
    tb.set_nerf_camera_matrix(fx["nerf_cams"][int(cam_i)][:3])
    np.testing.assert_allclose(tb.camera_matrix, fx["ngp_cams"][i], rtol=0, atol=1e-6)  # same ngp-convention matrix as the reference's m_camera
    tb.fov_axis = 0
    tb.fov = float(fx["fov_deg"])
    tb.snap_to_pixel_centers = bool(snap_px)
    tb.nerf.render_min_transmittance = float(min_t)
    tb.background_color = [r, g, b, a]
    tb.exposure = float(exposure)
    got = tb.render(want.shape[1], want.shape[0], int(spp), linear=bool(linear))
    p = assert_frames_agree(got, want, 70.0, f"classic {i}")
    print(f"classic {i}: PSNR vs reference {p:.1f} dB")
    tb.exposure = 0.0


def _pyngp_request(pyngp, fx, case, path):
    W = H = fx["bl_single"].shape[0]
    def masks(lst):
        out = []
        for m in lst:
            t, mode = np.asarray(m["transform"], np.float32), pyngp.MaskMode(m["mode"])
            if m["shape"] == 0:
                out.append(pyngp.Mask3D.Box(m["dims"], t, mode, m["feather"], m["opacity"]))
            elif m["shape"] == 1:
                out.append(pyngp.Mask3D.Cylinder(m["dims"][0], m["dims"][1], t, mode, m["feather"], m["opacity"]))
            else:
                out.append(pyngp.Mask3D.Sphere(m["dims"][0], t, mode, m["feather"], m["opacity"]))
        return out
    ds = pyngp.DownsampleInfo.MakeFromMip((W, H), case.get("mip", 0))
    out = pyngp.RenderOutputProperties((W, H), ds, 1, pyngp.ColorSpace(case.get("color_space", 0)), pyngp.TonemapCurve.Identity, case.get("exposure", 0.0),
                                       case.get("background", (0, 0, 0, 0)), bool(case.get("flip_y", 0)))
    sq = pyngp.SphericalQuadrilateralConfig(*case["sq"]) if "sq" in case else pyngp.SphericalQuadrilateralConfig.Zero()
    if "qh" in case:
        q = np.asarray(case["qh"], np.float32)
        qh = pyngp.QuadrilateralHexahedronConfig(pyngp.Quadrilateral3D(*q[:4]), pyngp.Quadrilateral3D(*q[4:]))
    else:
        qh = pyngp.QuadrilateralHexahedronConfig.Zero()
    cam = pyngp.RenderCameraProperties(fx["ngp_cams"][0], pyngp.CameraModel(case.get("model", 0)), float(fx["bl_focal"]), case.get("near", 0.0), case.get("aperture", 0.0),
                                       case.get("focus_z", 1.0), sq, qh)
    box = pyngp.BoundingBox([0, 0, 0], [1, 1, 1])
    nerfs = [pyngp.NerfDescriptor(path, box, np.asarray(n.get("transform", np.eye(4)), np.float32), pyngp.RenderModifiers(masks(n.get("masks", []))), n.get("opacity", 1.0))
             for n in case["nerfs"]]
    return pyngp.RenderRequest(out, cam, pyngp.RenderModifiers(masks(case.get("masks", []))), nerfs, pyngp.BoundingBox([-8, -8, -8], [8, 8, 8]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", BL_CASES)
def test_product_blender_render_matches_reference(fx, loaded_testbed, snapshot_path, name):
    """pyngp.Testbed.request_nerf_render_sync (K18) against Testbed::bl_render_frame of the reference: the NeRF is loaded from the reference-written snapshot
    through NerfDescriptor.snapshot_path, as the Blender add-on does."""
    import pyngp
    rq = _pyngp_request(pyngp, fx, _blender_cases(fx)[name], snapshot_path)
    got = loaded_testbed.request_nerf_render_sync(rq)
    p = assert_frames_agree(got, fx[f"bl_{name}"], 70.0, f"blender {name}")
    print(f"blender {name}: PSNR vs reference {p:.1f} dB")


@pytest.mark.gpu
def test_product_tonemap_curves(loaded_testbed, fx):
    """ETonemapCurve ACES / Hable / Reinhard (src/render_buffer.cu:272-329) through pyngp.Testbed.tonemap_curve: each curve applied by the CUDA path equals the
    curve's formula applied to the identity-curve frame (linear output, opaque background: the curve acts on the composited linear colour)."""
    tb = loaded_testbed
    tb.set_nerf_camera_matrix(fx["nerf_cams"][0][:3])
    tb.fov_axis = 0; tb.fov = float(fx["fov_deg"]); tb.snap_to_pixel_centers = True; tb.background_color = [0.2, 0.3, 0.4, 1.0]; tb.exposure = 0.0
    import pyngp
    tb.tonemap_curve = pyngp.TonemapCurve.Identity
    base = tb.render(64, 64, 1, linear=True)[..., :3].astype(np.float64)
    def rational(x, k):
        return (x * x * k[0] + k[1] * x + k[2]) / (k[3] * x * x + k[4] * x + k[5])
    aces = [0.6 * 0.6 * 2.51, 0.6 * 0.03, 0.0, 0.6 * 0.6 * 2.43, 0.6 * 0.59, 0.14]
    A, B, Cc, D, E, F = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
    k0, k1, k3, k4, k5 = A * F - A * E, Cc * B * F - B * E, A * F, B * F, D * F * F
    Wt = 11.2
    ws = (k3 * Wt * Wt + k4 * Wt + k5) / (k0 * Wt * Wt + k1 * Wt)
    hable = [4 * k0 * ws, 2 * k1 * ws, 0.0, 4 * k3, 2 * k4, k5]
    Y = base @ np.array([0.2126, 0.7152, 0.0722])
    want = {pyngp.TonemapCurve.ACES: rational(np.maximum(base, 0), aces), pyngp.TonemapCurve.Hable: rational(np.maximum(base, 0), hable),
            pyngp.TonemapCurve.Reinhard: np.maximum(base, 0) / (Y[..., None] + 1.0)}
    for curve, w in want.items():
        tb.tonemap_curve = curve
        got = tb.render(64, 64, 1, linear=True)[..., :3]
        np.testing.assert_allclose(got, w, rtol=2e-4, atol=2e-6, err_msg=str(curve))
    tb.tonemap_curve = pyngp.TonemapCurve.Identity
