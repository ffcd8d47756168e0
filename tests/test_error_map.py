"""K19, importance sampling of training pixels / images by accumulated error (nerf.training.sample_focal_plane_proportional_to_error /
sample_image_proportional_to_error): the CDF construction (construct_cdf_2d / construct_cdf_1d, src/testbed_nerf.cu:1984-2037, and the image normalisation of
:3000-3015), K1 drawing from the CDFs (sample_cdf_2d :991-1022, image_idx :1062-1083), K6 dividing the loss by the sampling density and depositing it into the
error map (:1448, :1465-1491), and the window cadence of Testbed::train_nerf (:2933-2939, :2971-3023).

Golden: tests/golden/ref_error_map.npz -- the reference's own kernels (built -fmad=false like the K1 golden) on the scene of ref_k1_nofma.npz and the seeded error
map of golden_inputs.error_map_inputs (oracle/gen_golden.py: error_map)."""
import ctypes as C
import os

import numpy as np
import pytest

from golden_inputs import ERROR_CDF_RES, ERROR_MAP_RES, error_map_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"xy": (True, False), "img": (False, True), "both": (True, True)}


def _golden(name):
    return np.load(os.path.join(GOLDEN, name))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _cdfs(orc, g, case):
    use_xy, use_img = CASES[case]
    return dict(cdf_x_cond_y=g["cdf_x_cond_y"] if use_xy else None, cdf_y=g["cdf_y"] if use_xy else None, cdf_img=g["cdf_img"] if use_img else None)


def _oracle_k1(orc, g, case):
    k1 = _golden("ref_k1_nofma.npz")
    imgs = orc.make_images(k1["images"], k1["xforms"], float(k1["fx"]), float(k1["fy"]))
    rng = orc.Pcg32(int(k1["rng_state"]), int(k1["rng_inc"]))
    with orc.error_sampling(**_cdfs(orc, g, case)):
        out = orc.generate_training_samples(int(k1["n_rays"]), k1["aabb"], int(k1["max_samples"]), rng, imgs, k1["bitfield"])
    return k1, imgs, rng, out


def _reslot_network_output(g, out1, max_samples):
    """The golden stores the network output per ray in ray-index order (concatenated); the oracle and the CUDA path keep rays in ray-index order too."""
    k = out1["n_kept"]
    assert np.array_equal(out1["ray_indices"][:k], g["both_ray_indices"]) and np.array_equal(out1["numsteps"][:k, 0], g["both_numsteps"])
    rgbsigma = np.zeros((max_samples, 4), np.float16)
    n_s = int(out1["counters"][0])
    assert n_s == g["k6_rgbsigma"].shape[0] and int(out1["numsteps"][0, 1]) == 0
    rgbsigma[:n_s] = g["k6_rgbsigma"]  # (bases grow with the ray index: the concatenation IS the sample order)
    return rgbsigma


def _ld_random_val_dim0(index, seed):
    """ld_random_val(index, seed, dim = 0), include/neural-graphics-primitives/random_val.cuh:159-268, in Python integers (Sobol dimension 0 = bit reversal)."""
    M = 0xFFFFFFFF

    def rev(x):
        return int(f"{x & M:032b}"[::-1], 2)

    def lk(x, s):
        x = (x + s) & M
        for c in (0x6c50b47c, 0xb82f1e52, 0xc7afe638, 0x8d22f6e6):
            x ^= (x * c) & M
        return x

    def scramble(x, s):
        return rev(lk(rev(x), s))
    index = scramble(index, seed)
    hc = (seed ^ ((0 + ((seed << 6) & M) + (seed >> 2)) & M)) & M
    return np.float32(scramble(rev(index), hc)) * np.float32(1.0 / (1 << 32))


# ---- CPU: the oracle against the reference's kernels ---------------------------------------------------------------------------------------------
def test_oracle_cdfs_match_reference(orc):
    """orc_construct_cdfs against construct_cdf_2d + construct_cdf_1d on the seeded error map: serial running sums, correctly rounded reciprocal, 1 % uniform
    blend -- every CDF entry and the un-normalised image sums bit-exact. The CDFs are non-decreasing and end at 1; the image probabilities (10 % uniform)
    sum to 1 and follow the image sums; rows / images without any error stay samplable."""
    g = _golden("ref_error_map.npz")
    em = error_map_inputs(g["cdf_x_cond_y"].shape[0])
    assert np.array_equal(em, g["error_map_in"]) and em.shape[1:] == (ERROR_CDF_RES[1], ERROR_CDF_RES[0])
    got = orc.construct_cdfs(em)
    assert np.array_equal(_bits(got["cdf_x_cond_y"]), _bits(g["cdf_x_cond_y"]))
    assert np.array_equal(_bits(got["cdf_y"]), _bits(g["cdf_y"]))
    assert np.array_equal(_bits(got["image_sums"]), _bits(g["image_sums"]))
    assert np.array_equal(_bits(got["cdf_img"]), _bits(g["cdf_img"]))
    for cdf in (got["cdf_x_cond_y"], got["cdf_y"], got["cdf_img"][None]):
        assert np.all(np.diff(cdf, axis=-1) > 0) and np.allclose(cdf[..., -1], 1.0, atol=2e-6)
    assert abs(float(got["pmf_img"].sum()) - 1.0) < 1e-5
    n = em.shape[0]
    want_pmf = 0.9 * em.sum((1, 2)) / em.sum() + 0.1 / n
    np.testing.assert_allclose(got["pmf_img"], want_pmf, rtol=1e-4)
    assert got["pmf_img"][1 % n] == pytest.approx(0.1 / n, rel=1e-3)  # the image without error keeps the uniform share
    # the empty row of every image keeps 1 % of a uniform row's mass
    rows = np.diff(np.concatenate([np.zeros((n, 1), np.float32), got["cdf_y"]], axis=1), axis=1)
    assert np.all(rows[[i for i in range(n) if i != 1 % n], 3] < 0.002) and np.all(rows > 0)


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_k1_with_cdfs_matches_reference(orc, case):
    """orc_generate_training_samples drawing pixels (xy), images (img) or both from the CDFs against the reference's generate_training_samples_nerf with the
    same CDFs: the kept rays, their origins / unnormalised directions bit for bit and their sample counts (the golden is stored in ray-index order)."""
    g = _golden("ref_error_map.npz")
    k1, imgs, rng, out = _oracle_k1(orc, g, case)
    k = out["n_kept"]
    assert k == int(g[f"{case}_ray_counter"]) and int(out["counters"][0]) == int(g[f"{case}_numsteps_counter"])
    assert np.array_equal(out["ray_indices"][:k], g[f"{case}_ray_indices"])
    assert np.array_equal(out["numsteps"][:k, 0], g[f"{case}_numsteps"])
    assert np.array_equal(_bits(out["rays"][:k]), _bits(g[f"{case}_rays"]))
    # the CDFs change the batch: not the uniform golden's ray set
    assert not np.array_equal(g[f"{case}_ray_indices"], np.sort(k1["ray_indices"][:int(k1["ray_counter"])]))


def test_oracle_image_sampling_follows_the_pmf(orc):
    """image_idx with the image CDF: over the 2048 rays of the batch the images are drawn with the frequencies of the pmf (the image with ten times the error
    gets the most rays, the one without error the 10 % / n floor), and the density handed to the loss is pmf x n."""
    g = _golden("ref_error_map.npz")
    n = g["cdf_img"].shape[0]
    pmf = np.diff(np.concatenate([[0.0], g["cdf_img"].astype(np.float64)]))
    cdf_img = np.ascontiguousarray(g["cdf_img"], np.float32)
    # ld_random_val(i, 0xdeadbeef) is a scrambled Sobol point set: binary-searching the CDF stratifies the images almost exactly
    u = np.array([_ld_random_val_dim0(i, 0xdeadbeef) for i in range(2048)], np.float32)
    img = np.minimum(np.searchsorted(cdf_img, u, side="left"), n - 1)
    freq = np.bincount(img, minlength=n) / 2048.0
    assert np.abs(freq - pmf).max() < 2e-3
    assert int(np.argmax(freq)) == 2 % n and int(np.argmin(freq)) == 1 % n
    # ... and these are the images the reference kernel drew: a kept ray starts at its camera's position
    k1 = _golden("ref_k1_nofma.npz")
    origins = np.stack([orc.effective_xform(x)[:, 3] for x in k1["xforms"]])
    for case in ("img", "both"):
        ro = g[f"{case}_rays"][:, :3]
        drawn = np.argmin(((ro[:, None, :] - origins[None]) ** 2).sum(-1), axis=1)
        assert np.array_equal(drawn, img[g[f"{case}_ray_indices"]])


def test_oracle_k6_error_map_matches_reference(orc):
    """orc_compute_loss with the CDFs and an error map against compute_loss_kernel_train_nerf on the same batch: compacted counts per ray exact, per-ray loss
    (divided by the sampling density) to 1e-3 of the largest, dL/dout (NOT divided) to 2e-3 of its range, and the deposited error map to 1e-3 of its largest
    texel (fp32 atomics in any order vs a serial sum; device __expf / powf vs libm). The loss differs from the density-free loss by exactly the density."""
    g = _golden("ref_error_map.npz")
    k1, imgs, rng, out1 = _oracle_k1(orc, g, "both")
    k, batch = out1["n_kept"], int(g["k6_batch"])
    rgbsigma = _reslot_network_output(g, out1, int(k1["max_samples"]))
    erx, ery = ERROR_MAP_RES
    em = np.zeros((len(k1["images"]), ery, erx), np.float32)
    with orc.error_sampling(error_map=em, **_cdfs(orc, g, "both")):
        out6 = orc.compute_loss(k, int(k1["n_rays"]), k1["aabb"], rng, batch, imgs, rgbsigma, out1["ray_indices"], out1["rays"], out1["numsteps"], out1["coords"],
                                float(g["k6_mean_density"][0]))
    assert out6["compacted"] == int(g["k6_compacted_counter"]) <= batch
    assert np.array_equal(out6["numsteps"][:k, 0], g["k6_compacted"])
    np.testing.assert_allclose(out6["loss"][:k], g["k6_loss"], rtol=0, atol=1e-3 * float(g["k6_loss"].max()))
    n_c = out6["compacted"]
    gd, wd = out6["dloss"][:n_c].astype(np.float32), g["k6_dloss"].astype(np.float32)
    assert gd.shape == wd.shape and np.abs(gd - wd).max() <= 2e-3 * np.abs(wd).max()
    assert g["k6_error_map"].max() > 0 and np.abs(em - g["k6_error_map"]).max() <= 1e-3 * float(g["k6_error_map"].max())
    # every ray with a compacted sample deposits exactly its loss x n_rays (bilinear weights sum to 1)
    assert float(em.sum()) == pytest.approx(float(out6["loss"][:k][out6["numsteps"][:k, 0] > 0].sum()) * int(k1["n_rays"]), rel=1e-4)
    # without the density division (CDFs only inside K1's pixel choice) the loss is larger / smaller by the density, the gradient is the same
    with orc.error_sampling(**_cdfs(orc, g, "both")):
        again = orc.compute_loss(k, int(k1["n_rays"]), k1["aabb"], rng, batch, imgs, rgbsigma, out1["ray_indices"], out1["rays"], out1["numsteps"], out1["coords"],
                                 float(g["k6_mean_density"][0]))
    assert np.array_equal(again["dloss"].view(np.uint16), out6["dloss"].view(np.uint16)) and np.array_equal(_bits(again["loss"]), _bits(out6["loss"]))


# ---- GPU: the CUDA path through the C ABI -----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def L():
    import torch
    import pyngp
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return pyngp.lib()


def _device_cdfs(g, case):
    import pyngp
    from gpu_util import dev
    use_xy, use_img = CASES[case]
    d_x, d_y, d_i = dev(g["cdf_x_cond_y"]), dev(g["cdf_y"]), dev(g["cdf_img"])
    c = pyngp.ErrorCdf(d_x.data_ptr() if use_xy else None, d_y.data_ptr() if use_xy else None, d_i.data_ptr() if use_img else None, ERROR_CDF_RES[0], ERROR_CDF_RES[1])
    return c, (d_x, d_y, d_i)


def _gpu_k1(L, k1, g, case):
    import torch
    import pyngp
    from gpu_util import dev, ptr, host, images_to_device, rng_struct
    scene = dict(images=k1["images"], xforms=k1["xforms"], fx=float(k1["fx"]), fy=float(k1["fy"]), cx=0.5, cy=0.5)
    meta, n_img, keep = images_to_device(scene)
    n_rays, max_samples = int(k1["n_rays"]), int(k1["max_samples"])
    aabb = np.ascontiguousarray(k1["aabb"], np.float32)
    d_bits = dev(k1["bitfield"])
    counters = torch.zeros(8, dtype=torch.int32, device="cuda"); ray_indices = torch.zeros(n_rays, dtype=torch.int32, device="cuda")
    rays = torch.zeros((n_rays, 6), dtype=torch.float32, device="cuda"); numsteps = torch.zeros((n_rays, 2), dtype=torch.int32, device="cuda")
    coords = torch.zeros((max_samples, 7), dtype=torch.float32, device="cuda")
    scratch = torch.zeros(int(L.ngpb_generate_training_samples_scratch_bytes(n_rays)), dtype=torch.uint8, device="cuda")
    rng = pyngp.Rng(int(k1["rng_state"]), int(k1["rng_inc"]))
    cdf, keep2 = (None, None) if case is None else _device_cdfs(g, case)
    pyngp.check(L.ngpb_generate_training_samples_cdf(None, n_rays, 0, n_rays, aabb.ctypes.data_as(C.c_void_p), max_samples, rng, n_img, ptr(meta), ptr(d_bits), 1, C.c_float(0.0),
                                                     ptr(counters), ptr(ray_indices), ptr(rays), ptr(numsteps), ptr(coords), ptr(scratch), None if cdf is None else C.byref(cdf)))
    return dict(counters=host(counters).view(np.uint32), ray_indices=host(ray_indices).view(np.uint32), rays=host(rays), numsteps=host(numsteps).view(np.uint32), coords=host(coords),
                dev=dict(meta=meta, n_img=n_img, keep=(keep, keep2), counters=counters, ray_indices=ray_indices, rays=rays, numsteps=numsteps, coords=coords, cdf=cdf, aabb=aabb, rng=rng))


@pytest.mark.gpu
def test_construct_error_cdfs_bit_exact(L, orc):
    """ngpb_construct_error_cdfs (row CDFs, row-sum CDFs, image CDF and probabilities, all on the device) against the reference's construct_cdf_2d /
    construct_cdf_1d outputs and the oracle's image normalisation: every float bit-exact; invalid arguments are refused."""
    import torch
    import pyngp
    from gpu_util import dev, ptr, host
    g = _golden("ref_error_map.npz")
    em = g["error_map_in"]
    n, ry, rx = em.shape
    want = orc.construct_cdfs(em)
    d_em = dev(em)
    d_x = torch.zeros((n, ry, rx), dtype=torch.float32, device="cuda"); d_y = torch.zeros((n, ry), dtype=torch.float32, device="cuda")
    d_i = torch.zeros(n, dtype=torch.float32, device="cuda"); d_p = torch.zeros(n, dtype=torch.float32, device="cuda")
    pyngp.check(L.ngpb_construct_error_cdfs(None, n, ry, rx, ptr(d_em), ptr(d_x), ptr(d_y), ptr(d_i), ptr(d_p)))
    assert np.array_equal(_bits(host(d_x)), _bits(g["cdf_x_cond_y"])) and np.array_equal(_bits(host(d_y)), _bits(g["cdf_y"]))
    assert np.array_equal(_bits(host(d_i)), _bits(g["cdf_img"])) and np.array_equal(_bits(host(d_p)), _bits(want["pmf_img"]))
    pyngp.check(L.ngpb_construct_error_cdfs(None, n, ry, rx, ptr(d_em), ptr(d_x), ptr(d_y), ptr(d_i), None))  # pmf is optional
    assert np.array_equal(_bits(host(d_i)), _bits(g["cdf_img"]))
    assert L.ngpb_construct_error_cdfs(None, n, ry, rx, None, ptr(d_x), ptr(d_y), ptr(d_i), None) != 0
    assert L.ngpb_construct_error_cdfs(None, 0, ry, rx, ptr(d_em), ptr(d_x), ptr(d_y), ptr(d_i), None) != 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_k1_with_cdfs_bit_exact(L, orc, case):
    """ngpb_generate_training_samples_cdf against the oracle (counters, ray indices, rays, per-ray counts and every sample record bit-exact) and against the
    reference kernel's kept rays; a null CDF struct / null members reproduce the uniform kernel bit for bit."""
    g = _golden("ref_error_map.npz")
    k1, imgs, rng, want = _oracle_k1(orc, g, case)
    got = _gpu_k1(L, k1, g, case)
    k, n_s = want["n_kept"], int(want["counters"][0])
    assert k > 100 and np.array_equal(got["counters"][:2], want["counters"])
    assert np.array_equal(got["ray_indices"][:k], want["ray_indices"][:k]) and np.array_equal(got["numsteps"][:k], want["numsteps"][:k])
    assert np.array_equal(_bits(got["rays"][:k]), _bits(want["rays"][:k]))
    assert np.array_equal(_bits(got["coords"][:n_s]), _bits(want["coords"][:n_s]))
    assert np.array_equal(got["ray_indices"][:k], g[f"{case}_ray_indices"]) and np.array_equal(_bits(got["rays"][:k]), _bits(g[f"{case}_rays"]))
    assert np.array_equal(got["numsteps"][:k, 0], g[f"{case}_numsteps"])
    if case == "both":
        plain = _gpu_k1(L, k1, g, None)
        kp = int(k1["ray_counter"])
        assert plain["counters"][1] == kp and np.array_equal(plain["ray_indices"][:kp], np.sort(k1["ray_indices"][:kp]))


@pytest.mark.gpu
def test_compute_loss_error_map_matches_reference_and_oracle(L, orc):
    """ngpb_compute_loss_error_map on the CUDA K1 output of the `both` batch: compaction identical to the oracle's, per-ray loss (divided by the density) to
    2e-4, dL/dout to 2e-3 of its range, the deposited error map against BOTH the oracle and the reference kernel (1e-3 of the largest texel). With null CDFs
    and no error map the entry point is ngpb_compute_loss bit for bit; an error map smaller than 2 x 2 is refused."""
    import torch
    import pyngp
    from gpu_util import dev, ptr, host
    g = _golden("ref_error_map.npz")
    k1, imgs, rng, out1 = _oracle_k1(orc, g, "both")
    got1 = _gpu_k1(L, k1, g, "both")
    d = got1["dev"]
    k, n_rays, batch = out1["n_kept"], int(k1["n_rays"]), int(g["k6_batch"])
    rgbsigma = _reslot_network_output(g, out1, int(k1["max_samples"]))
    erx, ery = ERROR_MAP_RES
    n_img = d["n_img"]
    em_want = np.zeros((n_img, ery, erx), np.float32)
    with orc.error_sampling(error_map=em_want, **_cdfs(orc, g, "both")):
        want = orc.compute_loss(k, n_rays, k1["aabb"], rng, batch, imgs, rgbsigma, out1["ray_indices"], out1["rays"], out1["numsteps"], out1["coords"], float(g["k6_mean_density"][0]))
    cfg = pyngp.LossConfig(128.0, (C.c_float * 3)(0, 0, 0), 1, 1, 0, 4, 2, 3, 1, 0.2)
    d_rgbsigma, d_mean = dev(rgbsigma), dev(np.asarray(g["k6_mean_density"], np.float32))
    scratch = torch.zeros(int(L.ngpb_compute_loss_scratch_bytes(n_rays)), dtype=torch.uint8, device="cuda")

    def run(cdf, with_map, res=(erx, ery), plain=False):
        numsteps = d["numsteps"].clone()
        coords_out = torch.zeros((batch, 7), dtype=torch.float32, device="cuda"); dloss = torch.zeros((batch, 4), dtype=torch.float16, device="cuda")
        loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda"); counters_out = torch.zeros(4, dtype=torch.int32, device="cuda")
        d_map = torch.zeros((n_img, res[1], res[0]), dtype=torch.float32, device="cuda") if with_map else None
        if plain:
            rc = L.ngpb_compute_loss(None, n_rays, d["aabb"].ctypes.data_as(C.c_void_p), d["rng"], batch, C.byref(cfg), n_img, ptr(d["meta"]), ptr(d["counters"]), ptr(d_rgbsigma),
                                     ptr(d["ray_indices"]), ptr(d["rays"]), ptr(numsteps), ptr(d["coords"]), ptr(d_mean), ptr(coords_out), ptr(dloss), ptr(loss), ptr(counters_out), ptr(scratch))
        else:
            rc = L.ngpb_compute_loss_error_map(None, n_rays, n_rays, d["aabb"].ctypes.data_as(C.c_void_p), d["rng"], batch, C.byref(cfg), n_img, ptr(d["meta"]), ptr(d["counters"]),
                                               ptr(d_rgbsigma), ptr(d["ray_indices"]), ptr(d["rays"]), ptr(numsteps), ptr(d["coords"]), ptr(d_mean), ptr(coords_out), ptr(dloss), ptr(loss),
                                               ptr(counters_out), ptr(scratch), None, None, None if cdf is None else C.byref(cdf), None if d_map is None else ptr(d_map), res[0], res[1])
        return rc, dict(total=int(host(counters_out).view(np.uint32)[0]), numsteps=host(numsteps).view(np.uint32)[:k].copy(), dloss=host(dloss).copy(), loss=host(loss)[:k].copy(),
                        em=None if d_map is None else host(d_map).copy())

    rc, got = run(d["cdf"], True)
    assert rc == 0
    assert got["total"] == want["compacted"] == int(g["k6_compacted_counter"]) and np.array_equal(got["numsteps"], want["numsteps"][:k])
    assert np.array_equal(got["numsteps"][:, 0], g["k6_compacted"])
    n_valid = min(got["total"], batch)
    gd, wd = got["dloss"][:n_valid].astype(np.float32), want["dloss"][:n_valid].astype(np.float32)
    assert np.abs(gd - wd).max() <= 2e-3 * np.abs(wd).max() + 1e-7
    assert np.abs(gd - g["k6_dloss"].astype(np.float32)).max() <= 2e-3 * np.abs(wd).max() + 1e-7
    np.testing.assert_allclose(got["loss"], want["loss"][:k], rtol=2e-4, atol=1e-9)
    np.testing.assert_allclose(got["loss"], g["k6_loss"], rtol=0, atol=1e-3 * float(g["k6_loss"].max()))
    scale = float(g["k6_error_map"].max())
    print(f"error map: max texel {scale:.4g}; vs oracle {np.abs(got['em'] - em_want).max() / scale:.2e}, vs reference {np.abs(got['em'] - g['k6_error_map']).max() / scale:.2e}")
    assert np.abs(got["em"] - em_want).max() <= 1e-3 * scale and np.abs(got["em"] - g["k6_error_map"]).max() <= 1e-3 * scale
    # null CDFs, no map: the plain kernel (the pixels K6 recovers are then the uniform ones -- a different batch, same code path as ngpb_compute_loss)
    rc_a, a = run(None, False)
    rc_b, b = run(None, False, plain=True)
    assert rc_a == 0 and rc_b == 0 and a["total"] == b["total"]
    assert np.array_equal(a["dloss"].view(np.uint16), b["dloss"].view(np.uint16)) and np.array_equal(_bits(a["loss"]), _bits(b["loss"]))
    assert not np.array_equal(_bits(a["loss"]), _bits(got["loss"]))
    rc, _ = run(d["cdf"], True, res=(1, 1))
    assert rc != 0


@pytest.mark.gpu
def test_testbed_error_map_sampling_follows_reference(tmp_path):
    """Testbed::train with the two switches against the reference's own Testbed on the same files (tests/golden/ref_error_map_train.npz; the small scene with
    one damaged image). Off by default, and then the window counters still follow the reference's cadence (CDFs due after 128 steps, then windows of 192,
    288, ... steps: src/testbed_nerf.cu:2971-3023; reset by reset_network, src/testbed.cu:2261-2264) but nothing is accumulated. Switched on: after each
    window the same window length and validity as the reference and the same error-map resolution -- exactly for the first window, within 30 % afterwards
    (it follows rays_per_batch, which the controller moves with the noise of the training); image probabilities sum to 1 with the 10 % / n floor, favour
    the damaged image (arg-max after the first window, as in the reference; above uniform after the second) and stay within 0.2 of the reference's (measured: 0.10,
    two trainings differ in the noise of their losses); the loss keeps falling; the run without the prefetched K1 of the next step behaves the same."""
    import pyngp
    import synthetic
    from golden_inputs import ERROR_SCENE, error_scene_images
    g = _golden("ref_error_map_train.npz")
    n, res, B, bad = ERROR_SCENE["n_images"], ERROR_SCENE["res"], ERROR_SCENE["batch"], ERROR_SCENE["damaged"]
    scene = dict(synthetic.make_lego_scene(n, res, device="cpu", seed=0))
    scene["images"] = error_scene_images(np.asarray(scene["images"]))
    tj = synthetic.write_transforms_json(scene, str(tmp_path))
    tb = pyngp.Testbed()
    tb.load_training_data(tj)
    tr = tb.nerf.training
    assert tr.sample_focal_plane_proportional_to_error is False and tr.sample_image_proportional_to_error is False
    assert tr.n_steps_between_error_map_updates == 128 and tr.n_steps_since_error_map_update == 0
    tb.train_n(130, B)
    assert tr.n_steps_between_error_map_updates == 192 and tr.n_steps_since_error_map_update == 2
    assert int(tb._get("error_cdf_valid")) == 0 and np.allclose(tr.get_error_map_pmf(), 1.0 / n)
    launches_off = tb.stats()["gpu_launches"]
    tb.reset()
    assert tr.n_steps_between_error_map_updates == 128 and tr.n_steps_since_error_map_update == 0
    tr.sample_focal_plane_proportional_to_error = True
    tr.sample_image_proportional_to_error = True
    tb.train_n(127, B)
    assert int(tb._get("error_cdf_valid")) == 0 and tr.n_steps_since_error_map_update == 127
    tb.train_n(1, B)
    losses, pmfs = [tb.loss], [tr.get_error_map_pmf()]
    state = lambda: [int(tb._get("error_map_res")), int(tb._get("error_map_res")), tr.n_steps_between_error_map_updates, int(tb._get("error_cdf_valid")), tr.n_steps_since_error_map_update]
    assert state() == g["states"][0].tolist()
    for k, w in enumerate(ERROR_SCENE["windows"][1:], 1):
        tb.train_n(w, B)
        assert state()[2:] == g["states"][k][2:].tolist(), (k, state(), g["states"][k])
        assert 0.7 * g["states"][k][0] <= state()[0] <= min(1.3 * g["states"][k][0], res), (k, state(), g["states"][k])
        losses.append(tb.loss); pmfs.append(tr.get_error_map_pmf())
    for k, pmf in enumerate(pmfs):
        ref = g["pmf"][k]
        print(f"window {k}: pmf ours {np.round(pmf, 4)} reference {np.round(ref, 4)} loss {losses[k]:.6f} (reference {float(g['losses'][k]):.6f})")
        assert abs(float(pmf.sum()) - 1.0) < 1e-4 and pmf.min() >= 0.1 / n - 1e-6
        assert np.abs(pmf - ref).max() < 0.2
        # (by the third window both networks explain the grey image through view dependence, and its share falls back to the others': reference 0.15, here 0.12 .. 0.23)
        assert pmf[bad] > 1.0 / n or k == 2
    assert int(np.argmax(pmfs[0])) == int(np.argmax(g["pmf"][0])) == bad and pmfs[0][bad] > 1.5 / n
    assert np.isfinite(losses[-1]) and losses[-1] < losses[0] and losses[-1] < 10.0 * float(g["losses"][-1]) + 1e-3
    # the same run without the prefetched sampling of the next step (fp32 atomics move the last bits of the weights only)
    tb2 = pyngp.Testbed()
    tb2.load_training_data(tj)
    tb2._set("overlap_sampling", 0)
    tb2.nerf.training.sample_focal_plane_proportional_to_error = True
    tb2.nerf.training.sample_image_proportional_to_error = True
    tb2.train_n(sum(ERROR_SCENE["windows"]), B)
    pmf2 = tb2.nerf.training.get_error_map_pmf()
    assert abs(float(pmf2.sum()) - 1.0) < 1e-4 and np.abs(pmf2 - pmfs[-1]).max() < 0.2
    assert tb2.nerf.training.n_steps_between_error_map_updates == 432 and np.isfinite(tb2.loss) and tb2.loss < losses[0]
    assert launches_off > 0
