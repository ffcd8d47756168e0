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
dist.destroy_process_group()
