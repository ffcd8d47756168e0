"""Two-or-more-GPU check of the NCCL data-parallel path (run under torchrun on a GPU box):
replicas must stay bit-identical, the loss must fall, and the step time is printed per rank."""
import os
import sys
import time

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
tb = pyngp.Testbed(device=local)
tb.init_data_parallel(rank, world)
tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
if "k19" in sys.argv[1:]:
    tb.nerf.training.sample_focal_plane_proportional_to_error = tb.nerf.training.sample_image_proportional_to_error = True
tb.train_n(40, 1 << 16)
l0 = tb.loss
t0 = time.perf_counter()
tb.train_n(200, 1 << 16)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
w, h, e = tb.get_params()
digest = torch.tensor([float(np.frombuffer(w.tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 40))], dtype=torch.float64, device="cuda")
all_d = [torch.zeros_like(digest) for _ in range(world)]
dist.all_gather(all_d, digest)
same = all(float(d) == float(all_d[0]) for d in all_d)
print(f"rank {rank}: loss {l0:.5f} -> {tb.loss:.5f}, {dt / 200 * 1e3:.3f} ms/step, rays/batch {tb.stats()['rays_per_batch']}, replicas identical: {same}", flush=True)
assert same and tb.loss < l0
if "k19" in sys.argv[1:]:
    # error-map importance sampling under data parallelism: every rank deposits its shard's rays, the maps are summed before the CDFs are built, so all
    # ranks must hold the same image probabilities and stay bit-identical while they draw their rays from them
    # (the switches were set before the first step: the CDFs were built after step 128 and the last 112 steps drew their rays from them)
    tr = tb.nerf.training
    pmf = torch.tensor(tr.get_error_map_pmf(), dtype=torch.float64, device="cuda")
    all_p = [torch.zeros_like(pmf) for _ in range(world)]
    dist.all_gather(all_p, pmf)
    w, h, e = tb.get_params()
    digest = torch.tensor([float(np.frombuffer(w.tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 40))], dtype=torch.float64, device="cuda")
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same_p = all(bool(torch.equal(p, all_p[0])) for p in all_p)
    same_w = all(float(d) == float(all_d[0]) for d in all_d)
    print(f"rank {rank}: error-map sampling: window {tr.n_steps_between_error_map_updates}, pmf sum {float(pmf.sum()):.5f} max {float(pmf.max()):.4f}, loss {tb.loss:.5f}, "
          f"pmf identical: {same_p}, replicas identical: {same_w}", flush=True)
    assert same_p and same_w and abs(float(pmf.sum()) - 1.0) < 1e-4 and np.isfinite(tb.loss)
dist.destroy_process_group()
