"""Per-image exposure optimisation (nerf.training.optimize_exposure): the exposure term of K6 (target colour x 2^exposure, src/testbed_nerf.cu:1403), the
exposure gradient it accumulates (:1558-1571) and the host-side Adam + zero-mean renormalisation that moves the exposures every 16 steps (:3105-3131).

Goldens come from the reference's own code (oracle/gen_golden.py: exposure_host, exposure_k6; oracle/gen_golden_full.py: exposure):
  ref_exposure_adam.npz   ngp::AdamOptimizer<Eigen::Array3f> driven through the reference's exposure block (host, the reference's Eigen)
  ref_k6_exposure.npz     compute_loss_kernel_train_nerf with non-zero exposures and an exposure-gradient buffer, on the K1 golden (B200, -fmad=false)
  ref_exposure_train.npz  ngp::Testbed trained with optimize_exposure on a dataset whose images carry known exposure errors"""
import ctypes as C
import os

import numpy as np
import pytest

from golden_inputs import (EXPOSURE_L2_REG, EXPOSURE_N_IMAGES, EXPOSURE_PER_CAMERA_LOSS_SCALE, EXPOSURE_SCENE, EXPOSURE_UPDATES, apply_image_exposures,
                           exposure_inputs, exposure_scene_offsets)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    return np.load(os.path.join(GOLDEN, name))


# ---- CPU --------------------------------------------------------------------------------------------------------------------------------------
def test_exposure_update_matches_reference_bit_exact():
    """ngpb_exposure_update (host) over 200 camera updates of 8 images against the reference's exposure block: Adam step on gradient * scale + l2 * exposure,
    then re-centring on a zero mean -- every exposure after every update bit-exact."""
    import pyngp
    L = pyngp.lib()
    d = exposure_inputs()
    want = _golden("ref_exposure_adam.npz")["exposures"]
    state = np.zeros((EXPOSURE_N_IMAGES, 10), np.float32)
    for u in range(EXPOSURE_UPDATES):
        g = np.ascontiguousarray(d["gradients"][u])
        L.ngpb_exposure_update(EXPOSURE_N_IMAGES, state.ctypes.data, g.ctypes.data, C.c_float(EXPOSURE_PER_CAMERA_LOSS_SCALE), C.c_float(EXPOSURE_L2_REG),
                               C.c_float(float(d["learning_rates"][u])))
        assert np.array_equal(state[:, 7:10].view(np.uint32), want[u].view(np.uint32)), f"update {u}"
        assert abs(float(state[:, 7:10].mean())) < 1e-6  # zero mean after the renormalisation
    assert np.abs(want[-1]).max() > 0.05  # the golden is not trivial


def _oracle_k6_on_golden(orc, exposure, linear, batch=None):
    """The oracle's K1 + K6 on the inputs of the K1 / K6 goldens (the reference's network output re-slotted per ray, as in test_oracle_cpu._k6_from_golden)."""
    k1 = _golden("ref_k1_nofma.npz"); k6 = _golden("ref_k6_nofma.npz")
    imgs = orc.make_images(k1["images"], k1["xforms"], float(k1["fx"]), float(k1["fy"]))
    rng = orc.Pcg32(int(k1["rng_state"]), int(k1["rng_inc"]))
    out1 = orc.generate_training_samples(int(k1["n_rays"]), k1["aabb"], int(k1["max_samples"]), rng, imgs, k1["bitfield"])
    n_kept = int(k1["ray_counter"])
    ref_slot_of_ray = {int(r): j for j, r in enumerate(k1["ray_indices"][:n_kept])}
    rgbsigma = np.zeros((int(k1["max_samples"]), 4), np.float16)
    for i in range(out1["n_kept"]):
        j = ref_slot_of_ray[int(out1["ray_indices"][i])]
        n, b_ref, b = int(out1["numsteps"][i, 0]), int(k1["numsteps"][j, 1]), int(out1["numsteps"][i, 1])
        rgbsigma[b: b + n] = k6["rgbsigma"][b_ref: b_ref + n]
    out6 = orc.compute_loss(out1["n_kept"], int(k1["n_rays"]), k1["aabb"], rng, int(batch or k6["batch"]), imgs, rgbsigma, out1["ray_indices"], out1["rays"],
                            out1["numsteps"], out1["coords"], float(k6["mean_density"][0]), linear_colors=bool(linear), exposure=exposure,
                            want_exposure_gradient=exposure is not None)
    return k1, k6, out1, out6, ref_slot_of_ray, rgbsigma, imgs, rng


def test_oracle_zero_exposure_is_the_plain_loss(orc):
    """exp(ln 2 * 0) = 1 exactly: with all-zero exposures the oracle's K6 reproduces its exposure-free output bit for bit."""
    k1 = _golden("ref_k1_nofma.npz")
    n_img = k1["images"].shape[0]
    a = _oracle_k6_on_golden(orc, None, 0)[3]
    b = _oracle_k6_on_golden(orc, np.zeros((n_img, 3), np.float32), 0)[3]
    assert a["compacted"] == b["compacted"] and np.array_equal(a["numsteps"], b["numsteps"])
    assert np.array_equal(a["dloss"].view(np.uint16), b["dloss"].view(np.uint16)) and np.array_equal(a["loss"].view(np.uint32), b["loss"].view(np.uint32))
    assert np.abs(b["exposure_gradient"]).max() > 0


@pytest.mark.parametrize("mode", ["srgb", "linear"])
def test_oracle_compute_loss_exposure_matches_reference(orc, mode):
    """orc_compute_loss_exposure against the reference kernel run with non-zero exposures: per-ray loss and dL/dout as in test_golden_compute_loss, and the
    accumulated exposure gradient per image and channel to 2e-3 of its largest component (device powf / __expf vs libm; fp32 atomics vs a serial sum)."""
    g = _golden("ref_k6_exposure.npz")
    batch = int(g["batch"])  # (large enough for every compacted sample: a clipped ray would drop out of the exposure gradient)
    k1, k6, out1, out6, ref_slot_of_ray, *_ = _oracle_k6_on_golden(orc, g["exposures"], mode == "linear", batch)
    assert out6["compacted"] == int(g[f"{mode}_compacted_counter"])
    gscale = np.abs(g[f"{mode}_dloss"].astype(np.float32)).max()
    worst_loss = worst_grad = 0.0
    for i in range(out1["n_kept"]):
        j = ref_slot_of_ray[int(out1["ray_indices"][i])]
        c_ref, b_ref = int(g[f"{mode}_numsteps_out"][j, 0]), int(g[f"{mode}_numsteps_out"][j, 1])
        c, b = int(out6["numsteps"][i, 0]), int(out6["numsteps"][i, 1])
        if b_ref + c_ref < batch and b + c < batch:
            assert c == c_ref
            worst_loss = max(worst_loss, abs(float(out6["loss"][i]) - float(g[f"{mode}_loss"][j])))
            if c:
                d = np.abs(out6["dloss"][b: b + c].astype(np.float32) - g[f"{mode}_dloss"][b_ref: b_ref + c].astype(np.float32)).max()
                worst_grad = max(worst_grad, float(d))
    assert worst_loss <= 1e-3 * float(g[f"{mode}_loss"].max())
    assert worst_grad <= 2e-3 * gscale
    # the exposures change the targets: the golden differs from the exposure-free one
    assert np.abs(g[f"{mode}_loss"] - k6["loss"]).max() > 1e-3 * float(k6["loss"].max()) or mode == "linear"
    want = g[f"{mode}_exposure_gradient"]
    assert np.abs(want).max() > 0
    assert out6["compacted"] <= batch
    assert np.abs(out6["exposure_gradient"] - want).max() <= 2e-3 * np.abs(want).max()


# ---- GPU --------------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def L():
    import torch
    import pyngp
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return pyngp.lib()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["srgb", "linear"])
def test_compute_loss_exposure_matches_reference_and_oracle(L, orc, mode):
    """ngpb_compute_loss_exposure on the inputs of the K1 golden with the golden's exposures: compaction identical to the oracle's, dL/dout to 2e-3 of range,
    per-ray loss to 2e-4, and the exposure gradient against BOTH the oracle and the reference kernel's own result (2e-3 of the largest component). With a null
    exposure pointer and with all-zero exposures the kernel reproduces ngpb_compute_loss bit for bit."""
    import torch
    import pyngp
    from gpu_util import dev, ptr, host, rng_struct, images_to_device
    g = _golden("ref_k6_exposure.npz")
    linear = mode == "linear"
    batch = int(g["batch"])
    k1, k6, out1, want, _, rgbsigma, imgs, rng = _oracle_k6_on_golden(orc, g["exposures"], linear, batch)
    n_rays, k = int(k1["n_rays"]), out1["n_kept"]
    scene = dict(images=k1["images"], xforms=k1["xforms"], fx=float(k1["fx"]), fy=float(k1["fy"]), cx=0.5, cy=0.5)
    meta, n_img, keep = images_to_device(scene)
    aabb = np.ascontiguousarray(k1["aabb"], np.float32)
    cfg = pyngp.LossConfig(128.0, (C.c_float * 3)(0, 0, 0), 1, 1, int(linear), 4, 2, 3, 1, 0.2)
    d_counters = dev(np.array([out1["counters"][0], k, 0, 0], np.uint32).view(np.int32))
    d_rgbsigma, d_ri, d_rays, d_coords = dev(rgbsigma), dev(out1["ray_indices"].view(np.int32)), dev(out1["rays"]), dev(out1["coords"])
    d_mean = dev(np.asarray(k6["mean_density"], np.float32))
    scratch = torch.zeros(int(L.ngpb_compute_loss_scratch_bytes(n_rays)), dtype=torch.uint8, device="cuda")

    def run(exposure, want_gradient):
        numsteps = dev(out1["numsteps"].view(np.int32))
        coords_out = torch.zeros((batch, 7), dtype=torch.float32, device="cuda"); dloss = torch.zeros((batch, 4), dtype=torch.float16, device="cuda")
        loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda"); counters_out = torch.zeros(4, dtype=torch.int32, device="cuda")
        d_e = None if exposure is None else dev(exposure)
        d_eg = torch.zeros((n_img, 3), dtype=torch.float32, device="cuda") if want_gradient else None
        pyngp.check(L.ngpb_compute_loss_exposure(None, n_rays, n_rays, aabb.ctypes.data_as(C.c_void_p), rng_struct(rng), batch, C.byref(cfg), n_img, ptr(meta), ptr(d_counters),
                                                 ptr(d_rgbsigma), ptr(d_ri), ptr(d_rays), ptr(numsteps), ptr(d_coords), ptr(d_mean), ptr(coords_out), ptr(dloss), ptr(loss),
                                                 ptr(counters_out), ptr(scratch), None if d_e is None else ptr(d_e), None if d_eg is None else ptr(d_eg)))
        return dict(total=int(host(counters_out).view(np.uint32)[0]), numsteps=host(numsteps).view(np.uint32)[:k].copy(), dloss=host(dloss).copy(), loss=host(loss)[:k].copy(),
                    eg=None if d_eg is None else host(d_eg).copy())

    got = run(g["exposures"], True)
    assert got["total"] == want["compacted"] and np.array_equal(got["numsteps"], want["numsteps"][:k])
    n_valid = min(got["total"], batch)
    gd, wd = got["dloss"][:n_valid].astype(np.float32), want["dloss"][:n_valid].astype(np.float32)
    assert np.abs(gd - wd).max() <= 2e-3 * np.abs(wd).max() + 1e-7
    np.testing.assert_allclose(got["loss"], want["loss"][:k], rtol=2e-4, atol=1e-9)
    ref = g[f"{mode}_exposure_gradient"]
    scale = np.abs(ref).max()
    print(f"exposure gradient ({mode}): |max| {scale:.4g}; vs oracle {np.abs(got['eg'] - want['exposure_gradient']).max() / scale:.2e}, vs reference {np.abs(got['eg'] - ref).max() / scale:.2e}")
    assert np.abs(got["eg"] - want["exposure_gradient"]).max() <= 2e-3 * scale
    assert np.abs(got["eg"] - ref).max() <= 2e-3 * scale
    # no exposures / zero exposures: the plain kernel
    plain, zero = run(None, False), run(np.zeros((n_img, 3), np.float32), False)
    assert plain["total"] == zero["total"] and np.array_equal(plain["dloss"].view(np.uint16), zero["dloss"].view(np.uint16))
    assert np.array_equal(plain["loss"].view(np.uint32), zero["loss"].view(np.uint32))
    assert not np.array_equal(plain["loss"].view(np.uint32), got["loss"].view(np.uint32))
    # an exposure gradient without exposures is refused
    eg = torch.zeros((n_img, 3), dtype=torch.float32, device="cuda")
    numsteps = dev(out1["numsteps"].view(np.int32))
    o = torch.zeros((batch, 7), dtype=torch.float32, device="cuda"); dl = torch.zeros((batch, 4), dtype=torch.float16, device="cuda"); co = torch.zeros(4, dtype=torch.int32, device="cuda")
    rc = L.ngpb_compute_loss_exposure(None, n_rays, n_rays, aabb.ctypes.data_as(C.c_void_p), rng_struct(rng), batch, C.byref(cfg), n_img, ptr(meta), ptr(d_counters), ptr(d_rgbsigma),
                                      ptr(d_ri), ptr(d_rays), ptr(numsteps), ptr(d_coords), ptr(d_mean), ptr(o), ptr(dl), None, ptr(co), ptr(scratch), None, ptr(eg))
    assert rc != 0


@pytest.mark.gpu
def test_optimize_exposure_follows_reference(tmp_path):
    """Whole loop against the reference's own Testbed: the small scene with known per-image exposure errors (golden_inputs.exposure_scene_offsets) baked into
    the images, trained for 2000 steps with nerf.training.optimize_exposure. The learned exposures must move towards undoing the errors (up to the common
    offset the renormalisation removes) at the reference's pace and correlate with what the reference learned on the same files
    (tests/golden/ref_exposure_train.npz). Also: off by default,
    nothing moves while the option is off, updates happen on the 16-step cadence, exposures stay zero-mean, reset_camera_extrinsics clears them."""
    import pyngp
    import synthetic
    g = _golden("ref_exposure_train.npz")
    n, res, B, steps = EXPOSURE_SCENE["n_images"], EXPOSURE_SCENE["res"], EXPOSURE_SCENE["batch"], EXPOSURE_SCENE["steps"]
    e = exposure_scene_offsets(n)
    assert np.array_equal(e, g["offsets"])
    scene = dict(synthetic.make_lego_scene(n, res, device="cpu", seed=0))
    scene["images"] = apply_image_exposures(np.asarray(scene["images"]), e)
    tj = synthetic.write_transforms_json(scene, str(tmp_path))
    order = sorted(range(n), key=lambda i: f"./train/r_{i}")  # both loaders sort the frames by file_path (nerf_loader.cu:356-358)
    assert np.array_equal(np.array(order), g["order"])
    expected = -(e[order] - e[order].mean(0))
    tb = pyngp.Testbed()
    tb.load_training_data(tj)
    tr = tb.nerf.training
    assert tr.optimize_exposure is False and tr.exposure_l2_reg == 0.0
    tb.train_n(20, B)
    assert np.all(tr.get_camera_exposures() == 0)
    tb.reset()
    tr.optimize_exposure = True
    tb.train_n(15, B)
    assert np.all(tr.get_camera_exposures() == 0) and tr.n_steps_since_cam_update == 15
    tb.train_n(1, B)
    first = tr.get_camera_exposures()
    assert np.count_nonzero(first) > first.size // 2 and tr.n_steps_since_cam_update == 0
    assert np.abs(first.mean(0)).max() < 1e-6
    # one Adam step at the network optimizer's learning rate moves a component by at most lr (before the mean is removed)
    assert np.abs(first).max() < 2.1e-2
    tb.train_n(steps - 16, B)
    got = tr.get_camera_exposures()
    ref = g["learned"][-1]
    rms = lambda a: float(np.sqrt(np.mean(np.square(a))))
    print(f"exposures after {tb.training_step} steps: rms(expected) {rms(expected):.4f}  rms(ours - expected) {rms(got - expected):.4f}  rms(reference - expected) {rms(ref - expected):.4f}"
          f"  rms(ours - reference) {rms(got - ref):.4f}  loss {tb.loss:.6f} (reference {float(g['losses'][-1]):.6f})")
    assert np.abs(got.mean(0)).max() < 1e-5
    # What the reference does with this dataset (measured, oracle/gen_golden_full.py:exposure): the exposure gradients are far below Adam's epsilon
    # (loss scale / rays x per-camera scale ~ 1e-9), so the exposures creep at ~2e-5 per step towards the expected values (slope 0.044 after 2000
    # steps, correlation 0.66) while the view-dependent colour of the network absorbs most of the error. Ours must do the same: same size, same
    # direction, and correlated with the reference's values (two runs of a chaotic training differ in the noise part of the gradients).
    slope = lambda a: float((a * expected).sum() / (expected ** 2).sum())
    print(f"slope towards the expected exposures: ours {slope(got):.4f}, reference {slope(ref):.4f}; rms ours {rms(got):.4f}, reference {rms(ref):.4f}; "
          f"corr(ours, reference) {float(np.corrcoef(got.ravel(), ref.ravel())[0, 1]):.3f}")
    assert 0.6 * rms(ref) < rms(got) < 1.6 * rms(ref)
    assert 0.6 * slope(ref) < slope(got) < 1.6 * slope(ref)
    assert float(np.corrcoef(got.ravel(), ref.ravel())[0, 1]) > 0.7
    assert np.isfinite(tb.loss) and tb.loss < 3.0 * float(g["losses"][-1]) + 1e-3
    # setting exposures by hand, then reset
    tr.set_camera_exposures(np.full((n, 3), 0.25, np.float32))
    assert np.all(tr.get_camera_exposures() == 0.25)
    tr.reset_camera_extrinsics()
    assert np.all(tr.get_camera_exposures() == 0)
