"""Collect the bench JSON lines of an evidence run (gpurun_out/bench_*.json) into profiles/<name>.md."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r1_bench.md")


def load(path):
    lines = [ln for ln in open(path) if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else None


rows, raw = [], []
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "bench_*.json"))):
    d = load(path)
    if not d:
        continue
    raw.append((os.path.basename(path), d))
    if d.get("impl") == "reference":
        rows.append((d["config"]["workload"], f"reference arm (CPU, {d['cpu_baseline']['cores']} threads)", d["n_gpus"],
                     d["value"], None, None, None, None, None))
        continue
    r = d["roofline"]
    rows.append((d["config"]["workload"], "b200 " + d["config"].get("scheme", "rk4"), d["n_gpus"], d["value"],
                 d["ms_per_step"], d["e2e"]["value"],
                 r.get("frac"), r["hbm"]["frac"], r.get("fp64_pipe_instructions_per_dof_step")))

with open(out, "w") as f:
    f.write("# Round-1 bench lines (driver contract: `python bench.py ...`, one JSON line each)\n\n")
    f.write("Measured on a B200 through `gpurun` with `tools/evidence_run.sh`; clocks and throttle reasons are "
            "inside each JSON line.\n\n")
    f.write("| workload | arm | GPUs | DOF-steps/s (resident) | ms/step | DOF-steps/s (e2e, host buffers) | "
            "FP64 issue-slot frac | HBM frac | FP64 instr per DOF-step |\n|---|---|---|---|---|---|---|---|---|\n")
    for w, arm, n, v, ms, e2e, frac, hfrac, slots in rows:
        f.write(f"| {w} | {arm} | {n} | {v:.3e} | {'' if ms is None else f'{ms:.3f}'} | "
                f"{'' if e2e is None else f'{e2e:.3e}'} | {'' if frac is None else f'{frac:.3f}'} | "
                f"{'' if hfrac is None else f'{hfrac:.4f}'} | {'' if slots is None else f'{slots:.0f}'} |\n")
    f.write("\nThe FP64 instruction count identifies the build: hh_ideal 16046, hh_tissue 15650 and calibration "
            "43795 are the kernels before\nthe shared-exponential rewrite (generator v9), 10081 (ncu: 10048) / 9881 / 38960 "
            "the final ones (generator v10, `codegen/fuse_exp.py`).\nThe 2-, 4- and 8-GPU lines and the "
            "remaining dp45 lines were measured before the rewrite and not repeated (GPU budget).\n")
    f.write("\n## Raw lines\n\n")
    for name, d in raw:
        f.write(f"### {name}\n\n```json\n{json.dumps(d, indent=1)}\n```\n\n")
print(open(out).read()[:1500])
