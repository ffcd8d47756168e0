"""K13 / K14, camera-extrinsics optimisation (nerf.training.optimize_extrinsics): gradients with respect to the network inputs, their reduction per
camera, the host-side per-camera Adam and the transform update.

tests/golden/ref_camera.npz holds the outputs of the reference's own code on the inputs of golden_inputs.camera_inputs (oracle/gen_golden.py:gen_camera):
kernel_grid's dy_dx + kernel_grid_backward_input, kernel_sh_backward, compute_cam_gradient_train_nerf (run on a B200) and AdamOptimizer<Vector3f> /
RotationAdamOptimizer / the update_transforms expression (host, the reference's Eigen)."""
import ctypes as C
import os

import numpy as np
import pytest

from golden_inputs import (CAM_AABB, CAM_ADAM_STEPS, CAM_N_IMAGES, CAM_N_KEPT, CAM_N_RAYS, N_CAM_SAMPLES, camera_inputs, grid_inputs)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_camera.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _rel(got, want):
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / max(float(np.abs(want).max()), 1e-30))


# ---- CPU: the oracle against the reference, the product's host functions against the reference -------------------------------------------------
def test_oracle_grid_input_gradient_matches_reference(orc, golden):
    """orc_grid_input_gradient against the reference kernels: the same trilinear-derivative terms, summed over the 32 features in the reference's order."""
    m = orc.model(aabb_scale=1)
    table = grid_inputs(m.n_grid_params)[0]
    d = camera_inputs()
    # the level scales as the device evaluates them (the reference computes them inside the kernel, grid.h:194-199)
    scales = np.load(os.path.join(os.path.dirname(GOLDEN), "ref_grid.npz"))["device_scales"]
    got = orc.grid_input_gradient(m, table, d["positions"], d["dL_dencoded"], scales=scales)
    assert _rel(got, golden["grid_dx"]) <= 2e-6


def test_oracle_sh_input_gradient_matches_reference(orc, golden):
    d = camera_inputs()
    got = orc.sh4_input_gradient(d["dirs"], d["dL_dsh"])
    assert _rel(got, golden["sh_dx"]) <= 2e-6


def test_oracle_cam_gradient_matches_reference(orc, golden):
    """orc_compute_cam_gradient (double accumulation) against the reference kernel's fp32 atomics: 1e-5 of the largest component."""
    d = camera_inputs()
    pos, rot = orc.compute_cam_gradient(CAM_N_KEPT, CAM_N_RAYS, CAM_N_IMAGES, CAM_AABB, d["ray_indices"], d["rays"], d["numsteps"], d["coords"], d["coords_gradient"])
    assert _rel(pos, golden["cam_pos_gradient"]) <= 1e-5
    assert _rel(rot, golden["cam_rot_gradient"]) <= 1e-5


def test_camera_adam_matches_reference_bit_exact(golden):
    """ngpb_camera_adam_step (host) over 400 steps against ngp::AdamOptimizer<Vector3f> and ngp::RotationAdamOptimizer: every intermediate variable bit-exact
    (same operation order as Eigen's fixed-size reductions; the rotation optimizer's double-precision bias correction included)."""
    import pyngp
    L = pyngp.lib()
    d = camera_inputs()
    for rot, key in ((0, "adam_pos"), (1, "adam_rot")):
        lr = d["adam_lr"] * (200 if rot else 1)
        state = np.zeros(10, np.float32)
        got = np.zeros((CAM_ADAM_STEPS, 3), np.float32)
        for i in range(CAM_ADAM_STEPS):
            g = np.ascontiguousarray(d["adam_gradients"][i])
            L.ngpb_camera_adam_step(state.ctypes.data, g.ctypes.data, float(np.float32(lr[i])), rot)
            got[i] = state[7:]
        assert state[0] == CAM_ADAM_STEPS
        assert np.array_equal(got, golden[key]), f"{key}: {np.abs(got - golden[key]).max():.3e}"
    assert np.abs(golden["adam_rot"]).max() > 1.0  # the rotation case does leave the small-angle regime


def test_apply_camera_offsets_matches_reference_bit_exact(golden):
    """ngpb_apply_camera_offsets against the statements of Training::update_transforms evaluated with the reference's Eigen (zero rotation: transform untouched)."""
    import pyngp
    L = pyngp.lib()
    d = camera_inputs()
    for k in range(6):
        out = np.zeros(12, np.float32)
        L.ngpb_apply_camera_offsets(np.ascontiguousarray(d["offsets_xforms"][k]).ctypes.data, np.ascontiguousarray(d["offsets_pos"][k]).ctypes.data,
                                    np.ascontiguousarray(d["offsets_rot"][k]).ctypes.data, out.ctypes.data)
        assert np.array_equal(out, golden["offsets_applied"][k])


def test_ngp_matrix_round_trip():
    import pyngp
    rs = np.random.RandomState(0)
    m = rs.randn(3, 4).astype(np.float32)
    back = pyngp.ngp_matrix_to_nerf(pyngp.nerf_matrix_to_ngp(m, 0.33, (0.5, 0.5, 0.5)), 0.33, (0.5, 0.5, 0.5))
    assert np.allclose(back, m, atol=1e-6)


# ---- GPU: the product kernels -------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def L():
    import pyngp
    lib = pyngp.lib()
    pyngp.check(lib.ngpb_check_device(0))
    return lib


def _grid_struct(L, aabb_scale=1):
    import pyngp
    return pyngp.grid_init(aabb_scale=aabb_scale, device_scales=True)[0]


@pytest.mark.gpu
def test_nerf_input_gradient_matches_reference(L, orc, golden):
    """ngpb_nerf_input_gradient against the reference's grid input gradient and kernel_sh_backward: 2e-6 of the largest component (the per-level terms are the
    reference's; the 16 levels are summed by a shuffle tree instead of sequentially)."""
    import pyngp
    from gpu_util import dev, ptr, host
    import torch
    m = orc.model(aabb_scale=1)
    table = grid_inputs(m.n_grid_params)[0]
    d = camera_inputs()
    n = N_CAM_SAMPLES
    coords = np.zeros((n, 7), np.float32)
    coords[:, :3] = d["positions"]; coords[:, 4:] = d["dirs"]
    out = torch.full((n, 7), 7.0, dtype=torch.float32, device="cuda")
    g = _grid_struct(L)
    pyngp.check(L.ngpb_nerf_input_gradient(None, C.byref(g), ptr(dev(table)), ptr(dev(coords)), n, ptr(dev(d["dL_dencoded"])), ptr(dev(d["dL_dsh"])), ptr(out)))
    got = host(out)
    assert _rel(got[:, :3], golden["grid_dx"]) <= 2e-6
    assert _rel(got[:, 4:], golden["sh_dx"]) <= 2e-6
    assert np.all(got[:, 3] == 0)
    # without SH gradients the direction part is zero; n = 0 is a no-op
    pyngp.check(L.ngpb_nerf_input_gradient(None, C.byref(g), ptr(dev(table)), ptr(dev(coords)), n, ptr(dev(d["dL_dencoded"])), None, ptr(out)))
    assert np.all(host(out)[:, 4:] == 0)
    pyngp.check(L.ngpb_nerf_input_gradient(None, C.byref(g), ptr(dev(table)), ptr(dev(coords)), 0, ptr(dev(d["dL_dencoded"])), None, ptr(out)))
    assert L.ngpb_nerf_input_gradient(None, C.byref(g), None, ptr(dev(coords)), n, ptr(dev(d["dL_dencoded"])), None, ptr(out)) != 0


@pytest.mark.gpu
def test_cam_gradient_matches_reference(L, golden):
    """ngpb_compute_cam_gradient against compute_cam_gradient_train_nerf: both accumulate with fp32 atomics in arbitrary order -> 1e-5 of the largest component.
    Rays beyond the kept-ray counter and rays without samples contribute nothing; the gradients accumulate across calls."""
    import pyngp
    from gpu_util import dev, ptr, host
    import torch
    d = camera_inputs()
    counter = dev(np.array([CAM_N_KEPT], np.uint32).view(np.int32))
    pos = torch.zeros((CAM_N_IMAGES, 3), dtype=torch.float32, device="cuda"); rot = torch.zeros_like(pos)
    args = (None, CAM_N_RAYS, CAM_N_RAYS, CAM_AABB.ctypes.data_as(C.c_void_p), ptr(counter), CAM_N_IMAGES, ptr(dev(d["ray_indices"].view(np.int32))), ptr(dev(d["rays"])),
            ptr(dev(d["numsteps"].view(np.int32))), ptr(dev(d["coords"])), ptr(dev(d["coords_gradient"])), ptr(pos), ptr(rot))
    pyngp.check(L.ngpb_compute_cam_gradient(*args))
    p1, r1 = host(pos).copy(), host(rot).copy()
    assert _rel(p1, golden["cam_pos_gradient"]) <= 1e-5
    assert _rel(r1, golden["cam_rot_gradient"]) <= 1e-5
    pyngp.check(L.ngpb_compute_cam_gradient(*args))
    assert _rel(host(pos), 2 * golden["cam_pos_gradient"]) <= 1e-5


@pytest.mark.gpu
def test_input_gradient_chain_matches_oracle(L, orc):
    """MLP backward with the SH-input gradient output (ngpb_nerf_mlp_forward_backward_sh) followed by ngpb_nerf_input_gradient against the oracle's
    orc_nerf_input_gradient on a random model: position and direction gradients within 2^-7 of their range for 99.9 % of the samples (fp16 hidden
    activations landing on the other side of a ReLU move single samples, as in test_mlp_forward_backward)."""
    import pyngp
    from gpu_util import dev, ptr, host
    import torch
    m = orc.model(aabb_scale=1)
    rs = np.random.RandomState(5)
    n = 128 * 40
    params = np.concatenate([(rs.uniform(-1, 1, 10240) * 0.25).astype(np.float16), (rs.randn(m.n_grid_params) * 0.3).astype(np.float16)])
    coords = rs.rand(n, 7).astype(np.float32)
    dout = (rs.randn(n, 4) * 0.05).astype(np.float16)
    g = _grid_struct(L)
    for l in range(16):
        m.scales[l] = g.scale[l]  # the oracle follows the device-evaluated level scales
    want = orc.nerf_input_gradient(m, params, coords, dout)
    d_params, d_coords = dev(params), dev(coords)
    enc = torch.zeros((n, 32), dtype=torch.float16, device="cuda")
    pyngp.check(L.ngpb_hash_encode_forward(None, C.byref(g), C.c_void_p(d_params.data_ptr() + 10240 * 2), ptr(d_coords), 7, n, ptr(enc)))
    ws = torch.zeros(int(L.ngpb_nerf_mlp_workspace_bytes()) // 4, dtype=torch.float32, device="cuda")
    denc = torch.zeros((n, 32), dtype=torch.float16, device="cuda"); dsh = torch.zeros((n, 16), dtype=torch.float16, device="cuda")
    grad = torch.zeros(10240, dtype=torch.float32, device="cuda")
    pyngp.check(L.ngpb_nerf_mlp_forward_backward_sh(None, ptr(d_params), ptr(enc), ptr(d_coords), ptr(dev(dout)), n, ptr(denc), ptr(grad), ptr(ws), ptr(dsh)))
    out = torch.zeros((n, 7), dtype=torch.float32, device="cuda")
    pyngp.check(L.ngpb_nerf_input_gradient(None, C.byref(g), C.c_void_p(d_params.data_ptr() + 10240 * 2), ptr(d_coords), n, ptr(denc), ptr(dsh), ptr(out)))
    got = host(out)
    for sl, what in ((slice(0, 3), "position"), (slice(4, 7), "direction")):
        err = np.abs(got[:, sl] - want[:, sl]) / np.abs(want[:, sl]).max()
        assert np.quantile(err, 0.999) <= 2.0 ** -7, f"{what}: 99.9th percentile {np.quantile(err, 0.999):.3e}"
        assert np.abs(want[:, sl]).max() > 0


def _pose_errors(tb, truth):
    pos, ang = [], []
    for i, t in enumerate(truth):
        m = tb.nerf.training.get_camera_extrinsics(i)
        pos.append(np.linalg.norm(m[:, 3] - t[:, 3]))
        r = m[:, :3] @ t[:, :3].T
        ang.append(np.degrees(np.arccos(np.clip((np.trace(r) - 1) / 2, -1, 1))))
    return float(np.mean(pos)), float(np.mean(ang))


@pytest.mark.gpu
def test_optimize_extrinsics_recovers_perturbed_cameras():
    """End to end (run.py --optimize_extrinsics flow, python_api.cu:811-844). A model is trained on the true cameras; then every training camera is replaced
    (set_camera_extrinsics) by one perturbed by 1.7 degrees / 0.024 scene units and the network is frozen (learning rate 0), so that only the per-camera
    offsets can lower the loss: the mean pose error against the true cameras must fall below half in rotation and below 0.6 in position. Also: the option's defaults,
    the update cadence (every n_steps_between_cam_updates steps, counted also while the option is off, testbed_nerf.cu:3026), reset_camera_extrinsics."""
    import pyngp
    import synthetic
    n_cam = 24
    scene = synthetic.make_lego_scene(n_cam, 96, device="cpu", seed=0)
    rs = np.random.RandomState(11)
    tb = pyngp.Testbed()
    tb.load_training_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
    tr = tb.nerf.training
    assert tr.optimize_extrinsics is False and tr.n_steps_between_cam_updates == 16
    assert abs(tr.extrinsic_learning_rate - 1e-3) < 1e-9 and abs(tr.extrinsic_l2_reg - 1e-4) < 1e-9
    truth = [tr.get_camera_extrinsics(i) for i in range(n_cam)]
    want0 = pyngp.ngp_matrix_to_nerf(np.asarray(scene["xforms"][0], np.float32).reshape(3, 4), tb._dataset_scale, tb._dataset_offset)
    assert np.allclose(truth[0], want0, atol=1e-6)
    tb.train_n(1000, 1 << 16)
    assert all(np.array_equal(tr.get_camera_extrinsics(i), truth[i]) for i in range(n_cam))  # option off: nothing moves
    for i, t in enumerate(truth):
        axis = rs.randn(3); axis /= np.linalg.norm(axis)
        ang = np.radians(1.7)
        K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        p = t.copy()
        p[:, :3] = (R @ t[:, :3]).astype(np.float32)
        p[:, 3] += (rs.randn(3) * 0.014).astype(np.float32)
        tr.set_camera_extrinsics(i, p)
    p0, a0 = _pose_errors(tb, truth)
    assert 1.6 < a0 < 1.8 and 0.015 < p0 < 0.035
    tb._set("learning_rate", 0.0)
    tr.optimize_extrinsics = True
    offsets = lambda: np.concatenate([np.concatenate(tr.get_camera_offsets(i)) for i in range(n_cam)])
    # 1000 steps ran since the last camera update, so the first step updates at once ...
    tb.train_n(1, 1 << 16)
    first = offsets()
    assert tr.n_steps_since_cam_update == 0 and np.count_nonzero(first) > 100
    # ... and then every 16th step
    tb.train_n(15, 1 << 16)
    assert np.array_equal(first, offsets()) and tr.n_steps_since_cam_update == 15
    tb.train_n(1, 1 << 16)
    assert not np.array_equal(first, offsets()) and tr.n_steps_since_cam_update == 0
    for _ in range(4):
        tb.train_n(800, 1 << 16)
        print(tb.training_step, _pose_errors(tb, truth), tb.loss)
    p1, a1 = _pose_errors(tb, truth)
    print(f"pose error: {p0:.4f} / {a0:.3f} deg -> {p1:.4f} / {a1:.3f} deg; loss {tb.loss:.6f}")
    assert np.isfinite(tb.loss)
    assert a1 < 0.5 * a0 and p1 < 0.6 * p0  # (measured 0.21 and 0.39)
    # reset: offsets zero again, transforms back at the (perturbed) dataset cameras
    tr.reset_camera_extrinsics()
    assert np.all(offsets() == 0)
    p2, a2 = _pose_errors(tb, truth)
    assert abs(a2 - a0) < 1e-4 and abs(p2 - p0) < 1e-6
    assert np.array_equal(tr.get_camera_extrinsics(1000), np.eye(4, dtype=np.float32)[:3])  # out of range: identity, as the reference returns
