"""Static SASS instruction mix of a generated model library's step kernel.

    python tools/sass_stats.py <libkem_*.so> [block]

Counts per kernel: FP64-pipe instructions (DFMA/DMUL/DADD/DSETP), everything
else, and the same restricted to the innermost loop (the RK4 stage loop:
between the last backward-branch target and that branch).  Used before
spending GPU time (B200_PROFILING.md: "check cuobjdump -sass here").
"""
import collections
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, table = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            addr, text = int(m.group(1), 16), m.group(2).strip()
            text = re.sub(r"^@!?U?P\w+\s+", "", text)
            table[cur].append((addr, text))
    return table


def summarize(insts):
    c = collections.Counter(t.split()[0].split(".")[0] for _, t in insts)
    dp = sum(v for k, v in c.items() if k in FP64)
    return c, dp, len(insts)


def inner_loop(insts):
    """Instructions of the smallest backward-branch loop that contains FP64 work."""
    best = None
    for addr, text in insts:
        m = re.match(r"BRA(?:\.\w+)*\s+.*?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [(a, t) for a, t in insts if tgt <= a <= addr]
                _, dp, n = summarize(body)
                if dp > 20 and (best is None or n < best[0]):
                    best = (n, body)
    return best[1] if best else []


def table(out_path):
    """profiles/fp64_slots.json: static FP64-pipe instruction counts of the builtin models."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "knp-emi-fenics-x_b200")]
    from knpemi_b200 import codegen
    from knpemi_b200.models import BUILTIN
    old = {}
    if os.path.exists(out_path):
        old = json.load(open(out_path))
    result = {}
    for name, mod in BUILTIN.items():
        path, em = codegen.model_library(mod)
        for kname, insts in kernels(path).items():
            if "kem_step_kernel" in kname and "ELi128E" in kname:
                c, dp, n = summarize(insts)
                lc, ldp, ln = summarize(inner_loop(insts))
                m_unroll = re.search(r"STAGE_UNROLL = (\d+);", em.source)
                stages_per_loop = int(m_unroll.group(1)) if m_unroll else 1
                iters = 4 // stages_per_loop                  # loop iterations per RK4 sub-step
                e = {"source_hash": em.source_hash, "instructions": n, "fp64": dp,
                     "loop_instructions": ln, "loop_fp64": ldp, "once_fp64": dp - ldp,
                     "stages_per_loop_iteration": stages_per_loop,
                     "static_per_dof_step_n_sub_25": 25 * iters * ldp + (dp - ldp)}
                prev = old.get(name, {})
                if prev.get("source_hash") == em.source_hash:       # keep ncu-measured figures
                    for k in ("ncu_per_dof_step", "ncu_n_sub", "ncu_dram_bytes_per_dof_step", "ncu_report"):
                        if k in prev:
                            e[k] = prev[k]
                result[name] = e
    with open(out_path, "w") as f:
        json.dump(result, f, indent=1, sort_keys=True)
    print(json.dumps(result, indent=1, sort_keys=True))


def main():
    if sys.argv[1] == "--table":
        return table(sys.argv[2])
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else None
    for name, insts in kernels(path).items():
        if "kem_step_kernel" not in name:
            continue
        m = re.search(r"ELi(\d+)E", name)
        block = m.group(1) if m else "?"
        if want and block != want:
            continue
        c, dp, n = summarize(insts)
        loop = inner_loop(insts)
        lc, ldp, ln = summarize(loop)
        print(f"block {block}: {n} instructions, {dp} FP64-pipe ({100.0 * dp / n:.0f} %)")
        print(f"  innermost FP64 loop: {ln} instructions, {ldp} FP64-pipe ({100.0 * ldp / max(ln, 1):.0f} %)")
        print("  loop mix:", ", ".join(f"{k} {v}" for k, v in lc.most_common(14)))


if __name__ == "__main__":
    main()
