"""Run under torchrun with 2 (or more) ranks on a GPU box: data-parallel training through the product path against single-GPU training of the same
global batch. Three runs from the same seed on the same scene, 60 steps each:
  A  world = N, fp32 gradient exchange (dp_half_gradients = 0)
  B  world = N, the default bf16 exchange
  C  world = 1 on rank 0 with batch N x B (the unsharded global batch)
Replicas must stay bit-identical within A and within B; the loss (read every 16 steps) of A and B must agree with C's within 15 % and with each other
within 5 % (the runs differ through per-shard roll-over padding, the sharded runs' larger inference budget, atomics order and, for B, bf16 rounding of
the partial gradients; this early in training the loss falls by a third every 16 steps, so a fraction of a step shows as several per cent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp
import synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = synthetic.make_lego_scene(16, 128, device=f"cuda:{local}", seed=0)
B, steps = 1 << 15, 64


def run(dp, half):
    tb = pyngp.Testbed(device=local)
    if dp:
        tb.init_data_parallel(rank, world)
        tb._set("dp_half_gradients", half)
    tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
    if not dp:  # start from the same GLOBAL ray count as the sharded runs (every rank starts at 4096 rays, testbed.h:374)
        import ctypes as C
        st = pyngp.TrainingState()
        pyngp.check(pyngp.lib().ngpb_testbed_get_training_state(tb._h, C.byref(st)))
        st.rays_per_batch *= world
        pyngp.check(pyngp.lib().ngpb_testbed_set_training_state(tb._h, C.byref(st)))
    losses = []
    for _ in range(steps // 16):
        tb.train_n(16, B if dp else B * world)
        losses.append(tb.loss)
    w, _, _ = tb.get_params()
    digest = torch.tensor([float(np.frombuffer(w.tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 40))], dtype=torch.float64, device="cuda")
    return losses, digest, w


def identical(digest):
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    return all(float(d) == float(all_d[0]) for d in all_d)


la, da, wa = run(True, 0)
same_a = identical(da)
lb, db, wb = run(True, 1)
same_b = identical(db)
dist.barrier()
ok = True
if rank == 0:
    lc, _, wc = run(False, 0)
    dev_a = max(abs(a - c) / c for a, c in zip(la, lc))
    dev_b = max(abs(b - c) / c for b, c in zip(lb, lc))
    print(f"loss every 16 steps: single GPU {['%.5f' % v for v in lc]}; fp32 exchange {['%.5f' % v for v in la]}; bf16 exchange {['%.5f' % v for v in lb]}")
    dev_ab = max(abs(a - b) / a for a, b in zip(la, lb))
    print(f"max relative loss deviation from the single-GPU run: fp32 exchange {dev_a:.4f}, bf16 exchange {dev_b:.4f}; bf16 vs fp32 exchange {dev_ab:.4f}; "
          f"replicas identical: {same_a} / {same_b}", flush=True)
    ok = same_a and same_b and dev_a <= 0.15 and dev_b <= 0.15 and dev_ab <= 0.05 and la[-1] < 0.6 * la[0]
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
