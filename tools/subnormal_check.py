"""Development aid: do the tcgen05 kind::f16 MMAs honour fp16 subnormal inputs? Backward of the plain MLP with loss gradients of a few 1e-6 (fp16 subnormals),
compared with the oracle (which rounds to fp16 including subnormals)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "blender-ngp_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import pyngp, oracle as orc
L = pyngp.lib()
rs = np.random.RandomState(0)
n = 128 * 8
w = (rs.uniform(-1, 1, 7168) * 0.25).astype(np.float16)
x = (rs.uniform(-1, 1, (n, 32)) * 0.5).astype(np.float16)
for mag in (1e-2, 1e-4, 1e-5, 3e-6, 1e-6):
    dy = np.zeros((n, 16), np.float16); dy[:, :3] = (rs.randn(n, 3) * mag).astype(np.float16)
    _, want_dx, want_g = orc.mlp_forward_backward(w, x, 2, dy)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dw, dx_, ddy = d(w), d(x), d(dy)
    denc = torch.zeros((n, 32), dtype=torch.float16, device="cuda"); grad = torch.zeros(7168, dtype=torch.float32, device="cuda")
    ws = torch.zeros(int(L.ngpb_nerf_mlp_workspace_bytes()), dtype=torch.uint8, device="cuda")
    pyngp.check(L.ngpb_mlp_forward_backward(None, C.c_void_p(dw.data_ptr()), C.c_void_p(dx_.data_ptr()), C.c_void_p(ddy.data_ptr()), n, C.c_void_p(denc.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(ws.data_ptr())))
    torch.cuda.synchronize()
    g = grad.cpu().numpy(); dxg = denc.cpu().numpy().astype(np.float32)
    print(f"|dy| ~ {mag:.0e}: |grad| gpu {np.abs(g).sum():.4e} oracle {np.abs(want_g).sum():.4e} ratio {np.abs(g).sum()/max(np.abs(want_g).sum(),1e-30):.3f};  |dx| gpu {np.abs(dxg).sum():.4e} oracle {np.abs(want_dx.astype(np.float32)).sum():.4e}; nonzero dy {np.count_nonzero(dy)}")
