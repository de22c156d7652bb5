"""Collect the bench JSON lines of an evidence run into profiles/<round>_bench.md.

    python tools/make_bench_report.py r2      # gpurun_out/r2f_bench_*.json -> profiles/r2_bench.md
"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
out = os.path.join(ROOT, "profiles", f"{rnd}_bench.md")


def load(path):
    lines = [ln for ln in open(path) if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def fmt(v, spec):
    return "" if v is None else format(v, spec)


rows, raw = [], []
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{rnd}f_bench_*.json"))):
    d = load(path)
    if not d:
        continue
    raw.append((os.path.basename(path), d))
    cfg = d["config"]
    if d.get("impl") == "reference":
        rows.append((cfg["workload"], f"reference arm (CPU port, {d['cpu_baseline']['cores']} threads)", d["n_gpus"],
                     d["value"], None, None, None, None, None, None, None))
        continue
    r, e = d["roofline"], d["e2e"]
    link = (e.get("link") or {}).get("frac")
    drop = e.get("dropin") or {}
    rows.append((cfg["workload"], "b200 " + cfg.get("scheme", "rk4"), d["n_gpus"], d["value"], d["ms_per_step"],
                 (d.get("sustained") or {}).get("value"), e["value"], link, r.get("frac"),
                 r.get("fp64_pipe_instructions_per_dof_step"),
                 "/".join(f"{drop[k]['value']:.2e}" for k in ("immediate", "deferred", "immediate_fresh_inputs") if k in drop)))

with open(out, "w") as f:
    f.write(f"# Round-{rnd[1:]} bench lines (driver contract: `python bench.py ...`, one JSON line each)\n\n")
    f.write("Measured on B200s through `gpurun` (`tools/evidence_r2_bench.sh`, `tools/n8_sweep_then_bench.sh`); clocks and "
            "throttle reasons are inside each JSON line.  *e2e / link floor* = time the exchange's own copies take with "
            "nothing else (measured in the same run) / time of the exchange.\n\n")
    f.write("| workload | arm | GPUs | DOF-steps/s (resident, K steps) | ms/step | resident, >= 1 s loop | DOF-steps/s (e2e, host buffers) | "
            "e2e / link floor | FP64 issue-slot frac | FP64 instr per DOF-step | unmodified call sequence: immediate / deferred / new inputs |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|\n")
    for w, arm, n, v, ms, sus, e2e, link, frac, slots, drop in rows:
        f.write(f"| {w} | {arm} | {n} | {v:.3e} | {fmt(ms, '.3f')} | {fmt(sus, '.3e')} | {fmt(e2e, '.3e')} | {fmt(link, '.2f')} | "
                f"{fmt(frac, '.3f')} | {fmt(slots, '.0f')} | {drop or ''} |\n")
    f.write("\n## Raw lines\n\n")
    for name, d in raw:
        f.write(f"### {name}\n\n```json\n{json.dumps(d, indent=1)}\n```\n\n")
print(open(out).read()[:2500])
