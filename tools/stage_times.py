"""Per-stage device times of the training step (serialised stages), for quick A/B of kernel variants selected by environment variables."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "blender-ngp_b200"))
import pyngp, synthetic
scene = synthetic.make_lego_scene(100, 800, device="cuda", as_numpy=True)
tb = pyngp.Testbed()
tb.load_training_images(list(scene["images"]), scene["xforms"], scene["fx"], scene["fy"])
tb.train_n(530)
tb._set("overlap_sampling", 0.0)
tb.profile_stages(True); tb.stage_times(reset=True)
tb.train_n(48)
st = tb.stage_times(reset=True)
print(" ".join(f"{k}={v[0] / max(v[1], 1) * 1e3:.0f}" for k, v in st.items() if v[1]), "| env", {k: v for k, v in os.environ.items() if k.startswith("NGPB_")})
