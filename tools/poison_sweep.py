"""Development aid: trains a few steps with one per-step buffer poisoned (0xFF) before every step (NGPB_POISON, csrc/testbed.cu) and reports whether the
result moved: a buffer whose poison changes the outcome is read before the step wrote it."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, "%s/blender-ngp_b200")
import pyngp, synthetic
scene = synthetic.make_lego_scene(8, 64, device="cpu", seed=0)
tb = pyngp.Testbed()
tb.load_training_images(scene["images"], scene["xforms"], scene["fx"], scene["fy"])
tb._set("reuse_encoding", float(sys.argv[1]))
tb.train_n(8, 1 << 14)
w, h, e = tb.get_params()
print(repr(float(np.abs(w.astype(np.float64)).sum())), bool(np.isfinite(w).all()), tb.loss)
''' % ROOT
names = ["encoded", "rgbsigma", "coords", "encoded_compacted", "coords_compacted", "dloss", "denc", "partials", "scratch", "loss", "rays/numsteps/ray_indices"]
for reuse in ("1", "0"):
    base = None
    for bit in [-1] + list(range(len(names))):
        env = dict(os.environ, NGPB_POISON=str(0 if bit < 0 else 1 << bit))
        out = subprocess.run([sys.executable, "-c", CHILD, reuse], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr.strip()[-200:]
        if bit < 0:
            base = line
        print(f"reuse={reuse} poison={'none' if bit < 0 else names[bit]:28s} {line}   {'' if bit < 0 or line.split()[0] == base.split()[0] else '<-- differs'}", flush=True)
