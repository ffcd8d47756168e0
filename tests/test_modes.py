"""ETestbedMode::Image and ETestbedMode::Sdf through pyngp.Testbed against the unmodified reference.

tests/golden/ref_modes.npz (oracle/gen_golden_full.py:gen_modes) holds what the reference's own Testbed produced on a B200 for the inputs of
golden_inputs.procedural_image / sdf_pool: initial parameters, the training batches of the first three steps, the loss read every 16th step over 1000
steps, and for the image model its trained parameters with the reference's inference, compute_image_mse and render of them."""
import hashlib
import os

import numpy as np
import pytest

from golden_inputs import MODE_IMAGE_BATCH, MODE_IMAGE_RES, MODE_SDF_BATCH, procedural_image, sdf_pool

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_modes.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def _psnr(a, b):
    return float(-10.0 * np.log10(max(float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)), 1e-20)))


# ---- CPU ----------------------------------------------------------------------------------------------------------------------------------------------
def test_mode_dispatch_and_config_mapping():
    """Testbed(mode) returns the mode's class without touching the GPU for Image; the network config maps onto ngpb_model_config with reset_network's derived
    per_level_scale (desired resolution = half the image's larger side) and the optimizer nesting of configs/image|sdf/base.json."""
    import pyngp
    from pyngp import modes
    tb = pyngp.Testbed(pyngp.TestbedMode.Image)
    assert isinstance(tb, modes.ImageTestbed) and tb.mode == pyngp.TestbedMode.Image and tb.training_step == 0
    with pytest.raises(RuntimeError):
        tb.train(1 << 14)  # no data
    with pytest.raises(RuntimeError):
        pyngp.Testbed(pyngp.TestbedMode.Volume)
    mc = modes.model_config_struct(modes.IMAGE_NETWORK_CONFIG, 2, 3, 256.0, 1337)
    assert (mc.n_pos_dims, mc.n_output_dims, mc.log2_hashmap_size, mc.base_resolution, mc.loss, mc.use_ema) == (2, 3, 24, 16, 0, 0)
    assert mc.per_level_scale == 0.0 and mc.desired_resolution == 256.0  # derived inside the library with the host's libm, as the reference does
    assert abs(mc.optimizer.learning_rate - 1e-2) < 1e-9 and mc.optimizer.decay_start == 20000 and mc.optimizer.decay_interval == 10000
    ms = modes.model_config_struct(modes.SDF_NETWORK_CONFIG, 3, 1, 2048.0, 1337)
    assert (ms.loss, ms.use_ema, ms.log2_hashmap_size) == (1, 1, 19) and abs(ms.optimizer.ema_decay - 0.95) < 1e-7 and abs(ms.optimizer.learning_rate - 1e-4) < 1e-10
    bad = dict(modes.IMAGE_NETWORK_CONFIG, network=dict(modes.IMAGE_NETWORK_CONFIG["network"], n_neurons=128))
    with pytest.raises(RuntimeError):
        modes.model_config_struct(bad, 2, 3, 256.0, 1337)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/sdf/bunny.obj"), reason="needs the reference's bundled mesh (build container only)")
def test_sdf_mesh_bounds_match_reference(golden):
    """SdfTestbed.load_training_data on the reference's bunny.obj: raw bounding box (inflated) and mesh_scale as Testbed::load_mesh computes them."""
    from pyngp import modes
    tb = object.__new__(modes.SdfTestbed)
    tb.sdf = modes._SdfNs()
    modes.SdfTestbed.load_training_data(tb, "/root/reference/data/sdf/bunny.obj")
    info = golden["sdf_bunny_mesh_info"]
    assert np.allclose(np.concatenate([tb._raw_aabb[0], tb._raw_aabb[1]]), info[:6], rtol=0, atol=2e-7 * np.abs(info[:6]).max())
    assert abs(tb.sdf.mesh_scale - info[6]) <= 2e-7 * info[6]


# ---- GPU ----------------------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def image_tb():
    import pyngp
    tb = pyngp.Testbed(pyngp.TestbedMode.Image)
    tb.load_image_data(procedural_image())
    return tb


@pytest.mark.gpu
def test_image_init_and_batches_bit_exact(image_tb, golden):
    """Same seed -> same model and same training data as the reference: the initial fp32 parameters (host xavier draws + device grid draws) and the positions
    of the first three training batches (m_rng stream, stratification, snapping) bit for bit; the targets (texel fetch of the device-converted image, linear ->
    sRGB) within one float ulp -- the reference's colour conversions are compiled with FMA contraction, this library rounds every operation (-fmad=false)."""
    tb = image_tb
    tb.reset(1337)
    assert tb.n_params == int(golden["image_n_params"])
    w, _, _ = tb.get_params()
    assert np.array_equal(w[:8192], golden["image_init_head"]) and np.array_equal(w[7168:7168 + 4096], golden["image_init_grid_head"])
    assert np.array_equal(_sha(w), golden["image_init_sha"])
    for step in range(3):
        tb.train(MODE_IMAGE_BATCH)
        pos, tgt = tb.training_batch(MODE_IMAGE_BATCH)
        assert np.array_equal(pos[:2048], golden[f"image_batch{step}_pos_head"]), f"step {step}: positions"
        assert np.array_equal(_sha(pos), golden[f"image_batch{step}_pos_sha"])
        assert np.abs(tgt[:2048] - golden[f"image_batch{step}_tgt_head"]).max() <= 1.2e-7, f"step {step}: targets"
        assert np.mean(tgt[:2048] == golden[f"image_batch{step}_tgt_head"]) > 0.6
    assert tb.training_step == 3


@pytest.mark.gpu
def test_image_training_follows_reference(image_tb, golden):
    """1000 steps at batch 2^16 next to the reference's run on the same batches: the first loss within 1e-3 (same parameters, same data; fp16 network) and the
    loss curve (every 16th step) within 20 % up to step 320, where the loss has fallen from 0.36 to 6e-6. Somewhere beyond that both runs turn noisy (learning rate 1e-2 at
    a loss of 5e-6: spikes of 10-100x in the reference's curve as well as in ours), so the tail is compared through its median, within a factor of 5."""
    tb = image_tb
    tb.reset(1337)
    ref = golden["image_loss_curve"]
    curve = []
    for step in range(int(golden["image_steps"])):
        tb.train(MODE_IMAGE_BATCH)
        if step % 16 == 0:
            curve.append(tb.loss)
    curve = np.array(curve[:len(ref)], np.float32)
    assert abs(curve[0] - ref[0]) <= 1e-3 * ref[0], (curve[0], ref[0])
    rel = np.abs(curve[1:21] - ref[1:21]) / ref[1:21]
    med, ref_med = float(np.median(curve[21:])), float(np.median(ref[21:]))
    print(f"image loss: ours {curve[0]:.5f} -> {curve[20]:.3e}, reference {ref[0]:.5f} -> {ref[20]:.3e}, max rel diff up to step 320: {rel.max():.3f}; tail median {med:.3e} vs {ref_med:.3e}")
    assert rel.max() <= 0.20  # (measured 0.115 - 0.116: the steep part around step 70, where the loss falls by 10x in 30 steps)
    assert 0.2 * ref_med <= med <= 5.0 * ref_med  # (one run in three of this round landed outside a factor of 2: the tail is the noise floor of an fp16 network at lr 1e-2)
    assert tb.compute_image_mse() < 5e-3  # (the last step may sit on a spike)


@pytest.mark.gpu
def test_image_inference_mse_render_on_reference_parameters(image_tb, golden):
    """The reference-trained parameters loaded into this model: inference at 4096 texel centres within one fp16 ulp of the reference's own, compute_image_mse
    (plain and byte-quantised; 8e-6, i.e. the size of the network's fp16 rounding) to 1 %, and two rendered frames (sRGB 1 spp with snapping; linear 2 spp, non-square) at >= 60 dB against the reference's
    render_frame output. Also a snapshot round trip of the model."""
    import pyngp
    tb = image_tb
    tb.set_params_half(golden["image_trained_params"])
    got = tb.inference(golden["image_query"])
    want = golden["image_inference"]
    assert np.abs(got - want).max() <= 2.0 ** -10 * max(1.0, float(np.abs(want).max())), np.abs(got - want).max()
    assert abs(tb.compute_image_mse() - float(golden["image_mse"])) <= 1e-2 * float(golden["image_mse"])
    assert abs(tb.compute_image_mse(True) - float(golden["image_mse_quantized"])) <= 1e-2 * float(golden["image_mse_quantized"])
    tb.snap_to_pixel_centers = True
    tb.background_color = [0.2, 0.3, 0.4, 1.0]
    for name in ("a", "b"):
        rw, rh, spp, linear = [int(v) for v in golden[f"image_render_{name}_cfg"]]
        fr = tb.render(rw, rh, spp, linear=bool(linear))
        ref = golden[f"image_render_{name}"].astype(np.float32)
        assert fr.shape == ref.shape
        p = _psnr(fr, ref)
        print(f"render {name}: {p:.1f} dB vs the reference frame, max abs {np.abs(fr - ref).max():.2e}")
        assert p >= 60.0
    tb.snap_to_pixel_centers = False


@pytest.mark.gpu
def test_image_snapshot_round_trip(image_tb, tmp_path):
    import pyngp
    tb = image_tb
    tb.reset(7)
    tb.train_n(40, 1 << 14)
    path = str(tmp_path / "image.msgpack")
    tb.save_snapshot(path)
    mse = tb.compute_image_mse()
    other = pyngp.Testbed(pyngp.TestbedMode.Image)
    other.load_image_data(procedural_image())
    other.load_snapshot(path)
    assert other.training_step == 40 and abs(other.compute_image_mse() - mse) <= 1e-6 * max(mse, 1e-9)
    # run.py's flow: constructor with a data path, frame() until a step count
    from PIL import Image as PILImage
    png = str(tmp_path / "img.png")
    PILImage.fromarray(procedural_image()).save(png)
    tb2 = pyngp.Testbed(pyngp.TestbedMode.Image, png, pyngp.modes.IMAGE_NETWORK_CONFIG)
    tb2.training_batch_size = 1 << 14
    while tb2.frame() and tb2.training_step < 50:
        pass
    assert tb2.training_step == 50 and 0 < tb2.loss < 0.2
    assert tb2.render(64, 48, 1, linear=False).shape == (48, 64, 4)


@pytest.mark.gpu
def test_sdf_mode_follows_reference(golden):
    """ETestbedMode::Sdf on a supplied pool (override_sdf_training_data semantics): initial parameters bit-exact, the shuffled batches of the first three
    steps bit-exact (tcnn shuffle with the step as seed), first loss within 1e-3, loss curve over 1000 steps within 15 % of the reference's after step 100."""
    import pyngp
    tb = pyngp.Testbed(pyngp.TestbedMode.Sdf)
    assert tb.n_params == int(golden["sdf_n_params"])
    w, _, _ = tb.get_params()
    assert np.array_equal(w[:8192], golden["sdf_init_head"]) and np.array_equal(_sha(w), golden["sdf_init_sha"])
    with pytest.raises(RuntimeError):
        tb.train(MODE_SDF_BATCH)  # no pairs yet
    pos, dist = sdf_pool()
    tb.set_unit_cube_pairs(pos, dist)
    ref = golden["sdf_loss_curve"]
    curve = []
    for step in range(int(golden["sdf_steps"])):
        tb.train(MODE_SDF_BATCH)
        if step < 3:
            bp, bd = tb.training_batch(MODE_SDF_BATCH)
            assert np.array_equal(bd[:1024, 0], golden[f"sdf_batch{step}_dist_head"])
            assert np.array_equal(_sha(bp), golden[f"sdf_batch{step}_pos_sha"]) and np.array_equal(_sha(bd), golden[f"sdf_batch{step}_dist_sha"])
        if step % 16 == 0:
            curve.append(tb.loss)
    curve = np.array(curve[:len(ref)], np.float32)
    assert abs(curve[0] - ref[0]) <= 1e-3 * ref[0], (curve[0], ref[0])
    rel = np.abs(curve[7:] - ref[7:]) / ref[7:]
    print(f"sdf loss: ours {curve[0]:.4f} -> {curve[-1]:.4f}, reference {ref[0]:.4f} -> {ref[-1]:.4f}, max rel diff after step 100: {rel.max():.3f}")
    assert rel.max() <= 0.15 and curve[-1] < 0.7 * curve[0]  # (measured 0.06 - 0.08 over several runs)
    # a pool smaller than the batch trains nothing (src/testbed_sdf.cu:1233), mesh-space pairs map through the loaded bounds
    step = tb.training_step
    tb.train(1 << 17)
    assert tb.training_step == step
    tb._raw_aabb = (np.array([-1, -1, -1], np.float32), np.array([1, 1, 1], np.float32)); tb.sdf.mesh_scale = 2.0
    tb.override_sdf_training_data(np.array([[0, 0, 0], [1, 1, 1]], np.float32).repeat(1 << 13, 0), np.array([0.5, -0.25], np.float32).repeat(1 << 13))
    tb.train(MODE_SDF_BATCH)
    bp, bd = tb.training_batch(MODE_SDF_BATCH)
    assert set(np.unique(bp)) == {0.5, 1.0} and set(np.unique(bd)) == {0.25, -0.125}
