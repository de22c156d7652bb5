"""Per-kernel share of device time from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.defaultdict(lambda: [0, 0.0, ""])
for r in rows:
    name, grid, ns = r[4], r[8], float(r[14].replace(",", ""))
    key = (name.split("(")[0][-48:], grid)
    tot[key][0] += 1
    tot[key][1] += ns
total = sum(v[1] for v in tot.values())
print(f"{len(rows)} launches, {total / 1e6:.3f} ms of device time")
for (name, grid), (cnt, ns, _) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {100 * ns / total:5.1f} %  {cnt:4d} x {ns / cnt / 1e6:9.4f} ms  grid {grid:>14s}  {name}")
