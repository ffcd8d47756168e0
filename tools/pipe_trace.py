import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo/blender-ngp_b200"); sys.path.insert(0, "/root/repo/tests")
import pyngp
L = pyngp.lib()
n = 148 * 6 * 3 * 128 + 148*128*2
rs = np.random.RandomState(0)
params = torch.from_numpy(((rs.rand(10240) - 0.5) * 0.5).astype(np.float16)).cuda()
enc = torch.from_numpy((rs.randn(n, 32) * 0.3).astype(np.float16)).cuda()
coords = torch.from_numpy(rs.rand(n, 7).astype(np.float32)).cuda()
out = torch.zeros((n, 4), dtype=torch.float16, device="cuda")
p = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    pyngp.check(L.ngpb_nerf_mlp_forward(None, p(params), p(enc), p(coords), n, p(out)))
torch.cuda.synchronize()
t = np.loadtxt(os.environ["NGPB_PIPE_TRACE"]).reshape(4, 6, 3, 8)
t0 = t[t > 0].min()
names = ["epi:mma_done seen", "epi:signalled", "iss:ready", "iss:committed"]
for s in (0, 1, 2):
    print("slot", s)
    for it in range(3):
        for step in range(5):
            print("  it", it, "step", step, " ".join(f"{names[r]}={int(t[r, s, it, step] - t0):6d}" for r in (2, 3, 0, 1)))
