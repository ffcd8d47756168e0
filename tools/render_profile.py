"""One classic 800 x 800 frame (K17) of the trained Lego-shaped scene inside a cudaProfilerStart/Stop window, for
  ncu --profile-from-start off --set full -k regex:render_ -o gpurun_out/render python tools/render_profile.py
Without a profiler it prints the frame time and the sample count."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp
import synthetic

scene = synthetic.make_lego_scene(100, 800, device="cuda", as_numpy=True)
tb = pyngp.Testbed()
tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
tb.train_n(530)
tb.camera_matrix = synthetic.nerf_matrix_to_ngp(synthetic.hemisphere_cameras(7, seed=3)[2])
tb.fov_axis = 0
tb.fov = math.degrees(synthetic.CAMERA_ANGLE_X)
tb.render(800, 800, 1, True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tb.render(800, 800, 1, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = [0.0] * 5
for k in range(5):
    tb.render(800, 800, 1, True)
    ms[k] = tb.last_render_ms
print(f"frame {sorted(ms)[2]:.3f} ms, {tb.last_render_samples} samples, {tb.last_render_samples / sorted(ms)[2] / 1e3:.1f} Msamples/s | env",
      {k: v for k, v in os.environ.items() if k.startswith("NGPB_")})
import numpy as np
_, bits = tb.get_density_grid()
b0 = bits[: 128 ** 3 // 8]
fine = np.unpackbits(b0).sum()
blocks = (b0.reshape(-1, 64).max(axis=1) > 0)
print(f"cascade 0: {fine} of {128 ** 3} cells occupied ({fine / 128 ** 3:.2%}); {blocks.sum()} of {blocks.size} 8^3 blocks non-empty ({blocks.mean():.2%})")
