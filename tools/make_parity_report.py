"""profiles/r2_parity_strict.md from the per-model reports the GPU test
tests/test_gpu_parity.py::test_strict_per_entry_parity_report leaves in gpurun_out/."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = [
    "# Strict per-entry parity of the CUDA path against the oracle (round 2, final build)",
    "",
    "`tests/diag_parity.py` / `tests/test_gpu_parity.py::test_strict_per_entry_parity_report` on a B200:",
    "20 000 DOFs x 10 PDE steps per model with the masked sticky stimulus, product build (`fast`) and",
    "triage build (`libm`: CUDA libm + IEEE division).  north_star's bar is 1e-10 relative on all states",
    "and I_ch.  *strict* = |got - want| / |want| per entry, no floor; *above* = entries whose strict error",
    "exceeds 1e-10; *floor needed* = the smallest f such that |got - want| / max(|want|, f colmax) < 1e-10",
    "everywhere.  The tests use f = 1e-6 for states and 3e-4 for currents.",
    "",
    "| model | build | column | strict max rel | above 1e-10 | max abs / colmax | floor needed |",
    "|---|---|---|---|---|---|---|",
]
files = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity_strict_*.json")))
if not files:
    sys.exit("no gpurun_out/parity_strict_*.json: run the GPU test first")
worst_state, worst_floor, n_above = 0.0, 0.0, 0
for path in files:
    rep = json.load(open(path))
    for math, cols in rep["builds"].items():
        for col, e in cols.items():
            out.append(f"| {rep['model']} | {math} | {col} | {e['strict_max_rel']:.1e} | {e['n_above_tol']} / {e['n']} | "
                       f"{e['abs_over_colmax']:.1e} | {e['min_floor_frac']:.1e} |")
            if col.startswith("state"):
                worst_state = max(worst_state, e["strict_max_rel"])
            else:
                worst_floor = max(worst_floor, e["min_floor_frac"])
                n_above = max(n_above, e["n_above_tol"])
out += ["",
        f"Every state entry passes strictly (worst {worst_state:.1e}).  Of the current entries at most {n_above} in 20 000",
        f"do not; the floor they need is at most {worst_floor:.1e} of the column maximum, in the `libm` build as well:",
        "a current is g (V - E) + pump terms, its error is the state error times the conductance (1e-14 .. 2e-13 of",
        "the column's scale at every DOF), and those entries are the DOFs where the terms cancel.",
        ""]
with open(os.path.join(ROOT, "profiles", "r2_parity_strict.md"), "w") as f:
    f.write("\n".join(out))
print("\n".join(out[-6:]))
