"""TEST INFRASTRUCTURE ONLY. Generates the golden vectors under tests/golden/ by running the REFERENCE's own kernels
(oracle/_ref/*.so, compiled from /root/reference by oracle/Makefile) on a B200:

    gpurun -- 'python oracle/gen_golden.py gpurun_out/golden'      # then copy gpurun_out/golden/*.npz to tests/golden/

Inputs are seeded numpy; outputs are the reference kernels' raw results. Nothing here is imported by the product.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def gen_grid(out_dir, aabb_scale=1, n=16384):
    import oracle
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_tcnn.so"))
    m = oracle.model(aabb_scale=aabb_scale)
    rs = np.random.RandomState(1234)
    table = (rs.randn(m.n_grid_params) * 0.5).astype(np.float16)
    positions = rs.rand(n, 3).astype(np.float32)
    positions[0] = 0.0; positions[1] = 1.0; positions[2] = [0.5, 0.25, 0.75]; positions[3] = [1.0, 0.0, 1.0]
    dy = (rs.randn(32, n) * 0.01).astype(np.float16)
    dirs = rs.rand(n, 3).astype(np.float32)
    offsets = (C.c_uint32 * 17)(*list(m.offsets[:17]))
    log2_pls = C.c_float(np.log2(np.float32(m.per_level_scale)))
    d_table, d_pos = dev(table), dev(positions)
    d_out = torch.zeros((32, n), dtype=torch.float16, device="cuda")
    assert lib.ref_grid_forward(n, 16, offsets, 16, log2_pls, ptr(d_table), ptr(d_pos), 3, ptr(d_out)) == 0
    d_grad = torch.zeros(m.n_grid_params, dtype=torch.float16, device="cuda")
    assert lib.ref_grid_backward(n, 16, offsets, 16, log2_pls, ptr(d_grad), ptr(d_pos), 3, ptr(dev(dy))) == 0
    d_sh = torch.zeros((n, 16), dtype=torch.float16, device="cuda")
    assert lib.ref_sh4(n, ptr(dev(dirs)), 3, ptr(d_sh), 16) == 0
    # the per-level scale as the DEVICE evaluates it (grid.h:194-199 runs inside the kernel)
    lvl = torch.arange(16, dtype=torch.float32, device="cuda")
    device_scales = (torch.exp2(lvl * float(log2_pls.value)) * 16 - 1.0)
    np.savez_compressed(os.path.join(out_dir, "ref_grid.npz"), aabb_scale=aabb_scale, table=table, positions=positions, encoded_soa=host(d_out),
                        dy_soa=dy, grad=host(d_grad), dirs=dirs, sh=host(d_sh), device_scales=host(device_scales), host_scales=np.array(m.scales[:16], np.float32))
    print("grid golden written; device/host scale bits equal:", np.array_equal(host(device_scales).view(np.uint32), np.array(m.scales[:16], np.float32).view(np.uint32)))


def gen_optimizer(out_dir):
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_tcnn.so"))
    rs = np.random.RandomState(77)
    n, n_matrix = 4096 + 20000, 4096
    w = (rs.randn(n) * 0.1).astype(np.float32)
    st = dict(w=dev(w), h=dev(w.astype(np.float16)), e=dev(np.zeros(n, np.float16)), m1=dev(np.zeros(n, np.float32)), m2=dev(np.zeros(n, np.float32)), s=dev(np.zeros(n, np.int32)))
    grads = []
    for step in range(1, 4):
        g = (rs.randn(n) * 5).astype(np.float16)
        g[n_matrix:][rs.rand(n - n_matrix) < 0.6] = 0
        grads.append(g)
        assert lib.ref_adam_step(n, n_matrix, C.c_float(128.0), C.c_float(1e-2), C.c_float(0.9), C.c_float(0.99), C.c_float(1e-15), C.c_float(1e-6),
                                 ptr(st["w"]), ptr(st["h"]), ptr(dev(g)), ptr(st["m1"]), ptr(st["m2"]), ptr(st["s"])) == 0
        old = 1 - 0.95 ** (step - 1); new = 1.0 / (1 - 0.95 ** step)
        assert lib.ref_ema_step(n, C.c_float(0.95), C.c_float(np.float32(old)), C.c_float(np.float32(new)), ptr(st["h"]), ptr(st["e"])) == 0
    np.savez_compressed(os.path.join(out_dir, "ref_optimizer.npz"), w0=w, grads=np.stack(grads), n_matrix=n_matrix,
                        w=host(st["w"]), h=host(st["h"]), e=host(st["e"]), m1=host(st["m1"]), m2=host(st["m2"]), s=host(st["s"]))
    print("optimizer golden written")


def gen_k1_k6(out_dir, libname, tag):
    import oracle
    import synthetic
    from conftest import scene_occupancy_bitfield
    path = os.path.join(HERE, "_ref", libname)
    if not os.path.exists(path):
        print(f"{libname} not built: skipping {tag}")
        return
    lib = C.CDLL(path)
    scene = synthetic.make_lego_scene(8, 64, device="cpu", seed=0)
    _, bits = scene_occupancy_bitfield(oracle)
    n_rays, max_samples, batch = 2048, 1 << 16, 1 << 14
    rng = oracle.pcg32(1337)
    aabb = np.array([0, 0, 0, 1, 1, 1], np.float32)
    xforms_cm = np.ascontiguousarray(np.stack([np.asarray(x, np.float32).reshape(3, 4).T.reshape(-1) for x in scene["xforms"]]))
    n_img = len(scene["images"])
    d_pix = dev(np.ascontiguousarray(scene["images"]))
    d_bits = dev(bits)
    ray_counter = torch.zeros(1, dtype=torch.int32, device="cuda"); numsteps_counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    ray_indices = torch.zeros(n_rays, dtype=torch.int32, device="cuda"); rays = torch.zeros((n_rays, 6), dtype=torch.float32, device="cuda")
    numsteps = torch.zeros((n_rays, 2), dtype=torch.int32, device="cuda"); coords = torch.zeros((max_samples, 7), dtype=torch.float32, device="cuda")
    st = lib.ref_generate_training_samples(n_rays, aabb.ctypes.data_as(C.c_void_p), max_samples, 0, C.c_uint64(rng.state), C.c_uint64(rng.inc),
                                           ptr(ray_counter), ptr(numsteps_counter), ptr(ray_indices), ptr(rays), ptr(numsteps), ptr(coords),
                                           n_img, 64, 64, C.c_float(scene["fx"]), C.c_float(scene["fy"]), C.c_float(0.5), C.c_float(0.5), ptr(d_pix),
                                           xforms_cm.ctypes.data_as(C.c_void_p), ptr(d_bits), 1, C.c_float(0.0))
    assert st == 0, st
    k1 = dict(ray_counter=int(host(ray_counter)[0]), numsteps_counter=int(host(numsteps_counter)[0]), ray_indices=host(ray_indices).view(np.uint32).copy(),
              rays=host(rays).copy(), numsteps=host(numsteps).view(np.uint32).copy(), coords=host(coords).copy())
    np.savez_compressed(os.path.join(out_dir, f"ref_k1_{tag}.npz"), images=np.ascontiguousarray(scene["images"]), xforms=scene["xforms"], fx=scene["fx"], fy=scene["fy"],
                        bitfield=bits, aabb=aabb, n_rays=n_rays, max_samples=max_samples, rng_state=np.uint64(rng.state), rng_inc=np.uint64(rng.inc), **k1)
    print(f"K1 golden ({tag}) written: {k1['ray_counter']} rays, {k1['numsteps_counter']} samples")

    # K6 on the reference's own K1 output with a seeded network output (padded_output_width = 16 as in the reference)
    n_s = k1["numsteps_counter"]
    rs = np.random.RandomState(4)
    net = np.zeros((max_samples, 16), np.float16)
    net[:n_s, :3] = rs.randn(n_s, 3).astype(np.float16)
    net[:n_s, 3] = (rs.randn(n_s) * 2.0 + 1.0).astype(np.float16)
    mean_density = np.array([0.005], np.float32)
    compacted_counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    coords_out = torch.zeros((batch, 7), dtype=torch.float32, device="cuda"); dloss = torch.zeros((batch, 16), dtype=torch.float16, device="cuda")
    loss = torch.zeros(n_rays, dtype=torch.float32, device="cuda")
    bg = np.zeros(3, np.float32)
    numsteps_io = numsteps.clone()
    st = lib.ref_compute_loss(n_rays, aabb.ctypes.data_as(C.c_void_p), 0, C.c_uint64(rng.state), C.c_uint64(rng.inc), batch, ptr(ray_counter), C.c_float(128.0), 16,
                              bg.ctypes.data_as(C.c_void_p), 1, 1, 0, n_img, 64, 64, C.c_float(scene["fx"]), C.c_float(scene["fy"]), C.c_float(0.5), C.c_float(0.5), ptr(d_pix),
                              xforms_cm.ctypes.data_as(C.c_void_p), ptr(dev(net)), ptr(compacted_counter), ptr(ray_indices), ptr(rays), ptr(numsteps_io),
                              ptr(coords), ptr(coords_out), ptr(dloss), 4, ptr(loss), 2, 3, 1, ptr(dev(mean_density)), C.c_float(0.2))
    assert st == 0, st
    np.savez_compressed(os.path.join(out_dir, f"ref_k6_{tag}.npz"), rgbsigma=net[:, :4].copy(), mean_density=mean_density, batch=batch,
                        compacted_counter=int(host(compacted_counter)[0]), numsteps_out=host(numsteps_io).view(np.uint32).copy(), coords_out=host(coords_out).copy(),
                        dloss=host(dloss)[:, :4].copy(), loss=host(loss).copy())
    print(f"K6 golden ({tag}) written: compacted {int(host(compacted_counter)[0])}")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for fn in (lambda: gen_grid(out), lambda: gen_optimizer(out), lambda: gen_k1_k6(out, "libref_ngp_nofma.so", "nofma"), lambda: gen_k1_k6(out, "libref_ngp.so", "fma")):
        try:
            fn()
        except Exception as e:  # keep going: each golden file is independent
            import traceback
            traceback.print_exc()
