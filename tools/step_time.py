"""Device time per training step (train_n, overlap on), for A/B of schedule variants selected by environment variables."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp, synthetic
scene = synthetic.make_lego_scene(100, 800, device="cuda", as_numpy=True)
tb = pyngp.Testbed()
tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
tb.train_n(530)
st = torch.cuda.ExternalStream(tb.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(st); tb.train_n(200); e1.record(st); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 200 * 1e3:.0f} us/step | env", {k: v for k, v in os.environ.items() if k.startswith("NGPB_")})
