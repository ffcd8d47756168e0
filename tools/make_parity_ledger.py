"""Regenerates PARITY.md: one row per test with the first paragraph of its docstring (what is compared, against what, at which tolerance)."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = (("tests/test_gpu_kernels.py", "GPU parity tests (`pytest -m gpu`)"), ("tests/test_oracle_cpu.py", "Oracle vs the reference's golden vectors and invariants (CPU)"),
         ("tests/test_abi.py", "ABI surface, containers, loader (CPU)"), ("tests/test_data_parallel.py", "Data parallelism: host logic (CPU, gloo, world size 2) and the two-GPU equivalence run"),
         ("tests/test_reference_fixtures.py", "Against fixtures produced by the unmodified reference Testbed (snapshot, frames, bitfield): oracle on CPU, product on GPU"),
         ("tests/test_camera_optimizer.py", "K13 / K14 camera-extrinsics optimisation against the reference's kernels and its host optimizers"),
         ("tests/test_modes.py", "Image and SDF Testbed modes against the reference's own modes"),
         ("tests/test_exposure.py", "Per-image exposure optimisation against the reference's kernel, its host optimizer and its Testbed"),
         ("tests/test_error_map.py", "K19 error-map importance sampling against the reference's kernels and its Testbed"),
         ("tests/test_online_dataset.py", "Datasets filled from arrays while training runs; Half / Float training images"))
out = ["# PARITY — what each test pins, against what, at which tolerance\n",
       "Generated from the test docstrings (`python tools/make_parity_ledger.py`). `gpu` tests call the CUDA path through the C ABI (`libngpb200.so`) and compare it with the",
       "CPU oracle (`oracle/`) and with golden vectors produced by the reference's own kernels on a B200 (`tests/golden/`, `oracle/gen_golden.py`); the other tests run",
       "without a GPU and pin the oracle itself, the ABI surface and the host logic.\n"]
for fn, title in FILES:
    tree = ast.parse(open(os.path.join(ROOT, fn)).read())
    out += [f"## {title} — `{fn}`\n", "| test | what it pins |", "|---|---|"]
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name.startswith("test_"):
            doc = " ".join((ast.get_docstring(node) or "").split()).replace("|", "\\|")
            params = ""
            for d in node.decorator_list:
                if isinstance(d, ast.Call) and getattr(d.func, "attr", "") == "parametrize":
                    try:
                        params = " ×" + str(len(ast.literal_eval(d.args[1])))
                    except Exception:
                        params = " (parametrized)"
            out.append(f"| `{node.name}`{params} | {doc} |")
    out.append("")
open(os.path.join(ROOT, "PARITY.md"), "w").write("\n".join(out) + "\n")
print("written PARITY.md")
