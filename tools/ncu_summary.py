"""Summarise an .ncu-rep (read here, on the CPU box) into the figures DESIGN.md / profiles/ cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_dofs_per_launch]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    path = sys.argv[1]
    n = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units, data = raw(path)
    col = {h: i for i, h in enumerate(hdr)}
    for r, row in enumerate(data):
        print(f"--- launch {r}: {row[col['Kernel Name']][:60]}")
        for k in KEYS:
            if k in col:
                print(f"  {k:72s} {row[col[k]]:>18s} {units[col[k]]}")
        stalls = sorted(((float(row[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', ''))
                         for h, i in col.items()
                         if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')
                         and row[i] not in ('', 'n/a')), reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("  warp-state samples: " + ", ".join(f"{nm} {100 * v / tot:.0f}%" for v, nm in stalls[:8]))
        def f(k):
            return float(row[col[k]].replace(',', ''))
        try:
            # per-thread instruction counters may be reported per-cycle; prefer derived sums if present
            dp_cycle = sum(f(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed")
                           for o in ("dfma", "dmul", "dadd"))
            cyc = f("smsp__cycles_active.avg")
            dp_total = dp_cycle * cyc
            print(f"  FP64 thread-instructions (dfma+dmul+dadd): {dp_total:.4e}"
                  + (f"  = {dp_total / n:.0f} per DOF-step" if n else ""))
        except (KeyError, ValueError):
            pass
        if n:
            try:
                b = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
                unit = units[col["dram__bytes_read.sum"]]
                mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
                print(f"  DRAM traffic: {b * mult / n:.1f} B per DOF-step")
            except (KeyError, ValueError):
                pass


if __name__ == "__main__":
    main()
