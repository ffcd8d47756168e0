"""Prints the tcgen05 / TMEM / mbarrier timing probe (csrc/umma_probe.cu) as a table: SM cycles on a B200. Usage: gpurun -- python tools/umma_probe.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp

NAMES = [f"A batch latency N={n} k={k} (issue k MMAs M128xN K16 + commit -> barrier seen)" for n in (64, 16) for k in (1, 2, 4, 8)]
for n in (64, 16):
    for k in (1, 2, 4):
        NAMES += [f"B pipelined batches N={n} k={k}: cycles per batch (16 batches, own barriers)", f"B   ... of which issue loop per batch"]
NAMES += ["C tcgen05.ld x32 + wait, one warp", "C 2 x (x32 + wait), four warps", "C per x32 load sustained, four warps",
          "D fence.proxy.async", "D tcgen05.fence::before_thread_sync", "D tcgen05.fence::after_thread_sync", "D 8 x STS.128 (tile layout) + proxy fence",
          "D mbarrier arrive + wait (self)", "D try_wait on a completed barrier", "E round trip: 2 MMAs + commit -> 4 warps ld 64 cols + fence + arrive -> issuer (per round)",
          "F cycles per MMA, 1 warp x 32 MMAs M128 N64 (uniform issue path)", "F cycles per 4 MMAs, 4 warps x 32 MMAs M128 N64", "F 1 warp N16", "F 4 warps N16 (per 4 MMAs)",
          "F 1 warp wgrad-shaped M64 N64 MN-major", "F 4 warps wgrad-shaped (per 4 MMAs)"]

out = (C.c_longlong * 64)()
sections = int(sys.argv[2]) if len(sys.argv) > 2 else 63
print('sections', sections, flush=True)
pyngp.check(pyngp.lib().ngpb_probe_umma(None, out, len(NAMES), sections))
res = {n: int(out[i]) for i, n in enumerate(NAMES)}
for n, v in res.items():
    print(f"{v:8d}  {n}")
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(res, f, indent=1)
