"""Warp-state samples and executed instructions by SASS opcode from an .ncu-rep (source page),
plus the shared-memory bank-conflict figure of the exp table look-ups.

    python tools/ncu_by_opcode.py gpurun_out/r2_prof_hh_ideal.ncu-rep > profiles/r2_ncu_hh_ideal_by_opcode.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True,
                     check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
col = {h: i for i, h in enumerate(hdr)}
samples, executed = collections.Counter(), collections.Counter()
lds_wave, lds_ideal = 0, 0
for r in rows:
    if len(r) != len(hdr) or r[0] == "Address":
        continue
    text = re.sub(r"^\s*(@!?U?P\w+\s+)?", "", r[col["Source"]])
    op = text.split()[0].split(".")[0] if text.split() else "?"
    samples[op] += int(r[col["# Samples"]] or 0)
    executed[op] += int(r[col["Instructions Executed"]] or 0)
    if op == "LDS":
        lds_wave += int(r[col["L1 Wavefronts Shared"]] or 0)
        lds_ideal += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
ts, te = sum(samples.values()) or 1, sum(executed.values()) or 1
print(f"# Warp-state samples and executed instructions by SASS opcode")
print(f"# from: ncu -i {sys.argv[1].split('/')[-1]} --page source --csv")
print(f"# {ts} samples, {te} warp-instructions executed")
print("opcode      samples%   executed%")
for op, n in executed.most_common(18):
    print(f"{op:10s} {100 * samples[op] / ts:6.1f}    {100 * n / te:6.1f}")
if lds_ideal:
    print(f"# LDS: {lds_wave} shared-memory wavefronts for {lds_ideal} ideal ({lds_wave / lds_ideal:.2f} x: bank conflicts of "
          "the per-thread exp-table index)")
