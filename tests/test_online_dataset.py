"""Datasets filled from arrays while training runs, and the non-8-bit image types they bring: Testbed.create_empty_nerf_dataset (python_api.cu:545,
src/testbed_nerf.cu:2635-2641), nerf.training.set_image (python_api.cu:56-76: float images, linear colours with premultiplied alpha -- EImageDataType::Float;
EXR files are kept as EImageDataType::Half), set_camera_intrinsics (:2502-2516), n_images_for_training (:2783-2799, :2897), read_rgba for the three data
types (common_device.cuh:677-705)."""
import ctypes as C

import numpy as np
import pytest


def linear_premultiplied(images_u8, dtype=np.float32):
    """RGBA8 sRGB with straight alpha -> linear colours times alpha, the representation read_rgba hands to the kernels for 8-bit images."""
    x = np.asarray(images_u8).astype(np.float64) / 255.0
    rgb = np.where(x[..., :3] <= 0.04045, x[..., :3] / 12.92, ((x[..., :3] + 0.055) / 1.055) ** 2.4)
    a = x[..., 3:4]
    return np.concatenate([rgb * a, a], axis=-1).astype(dtype)


def _k1_k6(orc, scene, images, n_rays=1024, batch=1 << 14, seed=3):
    from conftest import scene_occupancy_bitfield
    _, bits = scene_occupancy_bitfield(orc)
    imgs = orc.make_images(images, scene["xforms"], scene["fx"], scene["fy"])
    rng = orc.pcg32(seed)
    k1 = orc.generate_training_samples(n_rays, [0, 0, 0, 1, 1, 1], 1 << 16, rng, imgs, bits)
    n_s = int(k1["counters"][0])
    rs = np.random.RandomState(1)
    rgbsigma = np.zeros((1 << 16, 4), np.float16)
    rgbsigma[:n_s, :3] = rs.randn(n_s, 3).astype(np.float16)
    rgbsigma[:n_s, 3] = (rs.randn(n_s) * 2.0 + 1.0).astype(np.float16)
    k6 = orc.compute_loss(k1["n_kept"], n_rays, [0, 0, 0, 1, 1, 1], rng, batch, imgs, rgbsigma, k1["ray_indices"], k1["rays"], k1["numsteps"], k1["coords"], 0.005)
    return bits, rng, k1, rgbsigma, k6


# ---- CPU --------------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float16, 2e-3)])
def test_oracle_float_and_half_images_agree_with_bytes(orc, small_scene, dtype, tol):
    """read_rgba (common_device.cuh:677-705): an 8-bit image and the same image stored as linear premultiplied floats (Float) or halfs (Half) describe the same
    colours, so K1 keeps the same rays and K6 yields the same loss and gradients up to the rounding of the stored values. A negative red channel masks a
    pixel like 0x00FF00FF does in an 8-bit image."""
    images = np.asarray(small_scene["images"])
    _, _, k1_b, _, k6_b = _k1_k6(orc, small_scene, images)
    _, _, k1_f, _, k6_f = _k1_k6(orc, small_scene, linear_premultiplied(images, dtype))
    k = k1_b["n_kept"]
    assert k1_f["n_kept"] == k and np.array_equal(k1_f["ray_indices"][:k], k1_b["ray_indices"][:k]) and np.array_equal(k1_f["numsteps"][:k], k1_b["numsteps"][:k])
    assert k6_f["compacted"] == k6_b["compacted"]
    scale = float(np.abs(k6_b["loss"]).max())
    assert np.abs(k6_f["loss"] - k6_b["loss"]).max() <= tol * scale
    gs = float(np.abs(k6_b["dloss"].astype(np.float32)).max())
    assert np.abs(k6_f["dloss"].astype(np.float32) - k6_b["dloss"].astype(np.float32)).max() <= max(tol, 1e-3) * gs
    # masking: the same pixels masked in both representations drop the same rays
    masked_b = images.copy(); masked_b[:, ::2, :, :] = np.array([255, 0, 255, 0], np.uint8)  # 0x00FF00FF little-endian: R 255, G 0, B 255, A 0
    masked_f = linear_premultiplied(images, dtype); masked_f[:, ::2, :, 0] = -1.0
    _, _, m_b, _, _ = _k1_k6(orc, small_scene, masked_b)
    _, _, m_f, _, _ = _k1_k6(orc, small_scene, masked_f)
    assert 0 < m_b["n_kept"] < k and m_f["n_kept"] == m_b["n_kept"] and np.array_equal(m_f["ray_indices"][:m_b["n_kept"]], m_b["ray_indices"][:m_b["n_kept"]])


# ---- GPU --------------------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def L():
    import torch
    import pyngp
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return pyngp.lib()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_k1_k6_on_float_and_half_images(L, orc, small_scene, dtype):
    """K1 and K6 through the C ABI on Float / Half training images against the oracle on the same arrays: K1 bit-exact (incl. rays dropped by a negative red
    channel), K6 compaction exact, per-ray loss 2e-4, dL/dout 2e-3 of range."""
    import torch
    import pyngp
    from gpu_util import dev, ptr, host, rng_struct
    images = linear_premultiplied(np.asarray(small_scene["images"]), dtype)
    images[:, 5::7, :, 0] = -1.0  # some masked rows
    n_rays, batch = 1024, 1 << 14
    bits, rng, k1, rgbsigma, want = _k1_k6(orc, small_scene, images, n_rays, batch)
    n = len(images)
    pix = dev(images.view(np.uint8))
    arr = (pyngp.Image * n)()
    per = images[0].nbytes
    for i in range(n):
        arr[i].pixels = pix.data_ptr() + i * per
        arr[i].h, arr[i].w = images[i].shape[0], images[i].shape[1]
        arr[i].fx, arr[i].fy, arr[i].cx, arr[i].cy = small_scene["fx"], small_scene["fy"], 0.5, 0.5
        arr[i].image_type = pyngp.IMAGE_FLOAT if dtype == np.float32 else pyngp.IMAGE_HALF
        cm = np.asarray(small_scene["xforms"][i], dtype=np.float32).reshape(3, 4).T.reshape(-1).copy()
        eff = np.empty(12, np.float32)
        L.ngpb_effective_xform(cm.ctypes.data_as(C.c_void_p), eff.ctypes.data_as(C.c_void_p))
        for j in range(12):
            arr[i].raw_xform[j] = float(cm[j]); arr[i].xform[j] = float(eff[j])
    meta = dev(np.frombuffer(bytes(arr), dtype=np.uint8).copy())
    aabb = np.array([0, 0, 0, 1, 1, 1], np.float32)
    d_bits = dev(bits)
    counters = torch.zeros(8, dtype=torch.int32, device="cuda"); ray_indices = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
    rays = torch.zeros((n_rays, 6), dtype=torch.float32, device="cuda"); numsteps = torch.zeros((n_rays, 2), dtype=torch.int32, device="cuda")
    coords = torch.zeros((1 << 16, 7), dtype=torch.float32, device="cuda")
    scratch = torch.zeros(int(max(L.ngpb_generate_training_samples_scratch_bytes(n_rays), L.ngpb_compute_loss_scratch_bytes(n_rays))), dtype=torch.uint8, device="cuda")
    pyngp.check(L.ngpb_generate_training_samples(None, n_rays, aabb.ctypes.data_as(C.c_void_p), 1 << 16, rng_struct(rng), n, ptr(meta), ptr(d_bits), 1, C.c_float(0.0),
                                                 ptr(counters), ptr(ray_indices), ptr(rays), ptr(numsteps), ptr(coords), ptr(scratch)))
    k, n_s = k1["n_kept"], int(k1["counters"][0])
    assert k > 50 and np.array_equal(host(counters).view(np.uint32)[:2], k1["counters"])
    assert np.array_equal(host(ray_indices).view(np.uint32)[:k], k1["ray_indices"][:k]) and np.array_equal(host(numsteps).view(np.uint32)[:k], k1["numsteps"][:k])
    assert np.array_equal(host(coords)[:n_s].view(np.uint32), k1["coords"][:n_s].view(np.uint32))
    cfg = pyngp.LossConfig(128.0, (C.c_float * 3)(0, 0, 0), 1, 1, 0, 4, 2, 3, 1, 0.2)
    coords_out = torch.zeros((batch, 7), dtype=torch.float32, device="cuda"); dloss = torch.zeros((batch, 4), dtype=torch.float16, device="cuda")
    loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda"); counters_out = torch.zeros(4, dtype=torch.int32, device="cuda")
    pyngp.check(L.ngpb_compute_loss(None, n_rays, aabb.ctypes.data_as(C.c_void_p), rng_struct(rng), batch, C.byref(cfg), n, ptr(meta), ptr(counters), ptr(dev(rgbsigma)),
                                    ptr(ray_indices), ptr(rays), ptr(numsteps), ptr(coords), ptr(dev(np.array([0.005], np.float32))), ptr(coords_out), ptr(dloss), ptr(loss),
                                    ptr(counters_out), ptr(scratch)))
    assert int(host(counters_out).view(np.uint32)[0]) == want["compacted"] and np.array_equal(host(numsteps).view(np.uint32)[:k], want["numsteps"][:k])
    n_valid = min(want["compacted"], batch)
    g, w = host(dloss)[:n_valid].astype(np.float32), want["dloss"][:n_valid].astype(np.float32)
    assert np.abs(g - w).max() <= 2e-3 * np.abs(w).max() + 1e-7
    np.testing.assert_allclose(host(loss)[:k], want["loss"][:k], rtol=2e-4, atol=1e-9)


@pytest.mark.gpu
def test_dataset_filled_from_arrays_while_training():
    """The reference's online-dataset flow (NeRF-SLAM style): create_empty_nerf_dataset -> train() is a no-op while n_images_for_training = 0 -> set_image
    (float32 linear premultiplied, as the reference requires) + set_camera_intrinsics + set_camera_extrinsics for the first half -> n_images_for_training = n / 2
    -> training runs on those -> the rest is added, n_images_for_training = n (the occupancy grid is re-marked) -> the result matches a Testbed that loaded
    the same scene as 8-bit images in one go. Invalid frames and unset images are refused."""
    import pyngp
    import synthetic
    n, res, B = 8, 64, 1 << 14
    scene = synthetic.make_lego_scene(n, res, device="cpu", seed=0)
    images_f = linear_premultiplied(np.asarray(scene["images"]))
    tb = pyngp.Testbed()
    tb.create_empty_nerf_dataset(n, aabb_scale=1)
    tr = tb.nerf.training
    assert tr.n_images_for_training == 0
    tb.train_n(3, B)
    assert tb.training_step == 0  # nothing to train on (src/testbed_nerf.cu:2897)
    with pytest.raises(RuntimeError, match="Invalid frame index"):
        tr.set_image(n, images_f[0])
    with pytest.raises(RuntimeError):
        tr.set_image(0, images_f[0][..., :3])

    def add(i):
        tr.set_image(i, images_f[i], np.zeros((0, 0), np.float32), 1.0)
        tr.set_camera_intrinsics(i, fx=scene["fx"], fy=scene["fy"], cx=res / 2, cy=res / 2)
        tr.set_camera_extrinsics(i, np.asarray(scene["xforms"][i], np.float32).reshape(3, 4), convert_to_ngp=False)
    for i in range(n // 2):
        add(i)
    tr.n_images_for_training = n // 2 + 1
    with pytest.raises(RuntimeError, match="has not been set"):
        tb.train_n(1, B)
    tr.n_images_for_training = n // 2
    tb.train_n(60, B)
    assert tb.training_step == 60 and np.isfinite(tb.loss) and tb.loss > 0
    for i in range(n // 2, n):
        add(i)
    tr.n_images_for_training = n
    tb.train_n(240, B)
    assert tb.training_step == 300
    ref = pyngp.Testbed()
    ref.load_training_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
    ref.train_n(300, B)
    print(f"online dataset (float images): loss {tb.loss:.6f} after 60 + 240 steps; bulk 8-bit load: {ref.loss:.6f} after 300")
    assert np.isfinite(tb.loss) and tb.loss < 10.0 * ref.loss + 1e-3  # (measured 2.2e-4 vs 1.1e-4: the first 60 steps saw half of the images)
    # both render the scene: the held-in view agrees
    cam = np.asarray(scene["xforms"][0], np.float32).reshape(3, 4)  # (the training transform itself: the two sessions hold different NeRF -> ngp conversions)
    tb.camera_matrix = cam; ref.camera_matrix = cam
    a, b = tb.render(res, res, 1, True), ref.render(res, res, 1, True)
    mse = float(np.mean((a[..., :3] - b[..., :3]) ** 2))
    print(f"PSNR between the two models' renders of training view 0: {-10 * np.log10(mse + 1e-12):.1f} dB")
    assert -10 * np.log10(mse + 1e-12) > 22.0  # (two independently trained 300-step models of an 8-image scene; measured 29.9 dB)
