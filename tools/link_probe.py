"""Host-link ceilings of one B200 (pinned copies, both directions, several copy sizes), with
the host side cache-resident (one buffer reused: the PCIe link alone) and streaming through
DRAM (a span far above the last-level cache: what the exchange of a real PDE step sees).

    python tools/link_probe.py [dev]

Prints one JSON object per case: what `bench.py` divides the end-to-end exchange by."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "knp-emi-fenics-x_b200"))
from knpemi_b200 import _cabi  # noqa: E402

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for span_mb in (0, 960):
    for mb in (1, 5, 16, 80):
        reps = max(6, 480 // mb)
        r = _cabi.link_ceiling(dev, mb << 20, reps=reps, span=span_mb << 20)
        print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}), flush=True)
# asymmetric loads: what the exchange of one PDE step looks like (5 columns in, 3 or 4 out)
for span_mb in (0, 960):
    for rin, rout in ((80, 48), (80, 64), (112, 64)):
        h, d = _cabi.link_probe(dev, 5 << 20, rin, rout, span_mb << 20)
        print(json.dumps({"copy_bytes": 5 << 20, "span_mb": span_mb, "reps_h2d": rin, "reps_d2h": rout,
                          "h2d": round(h, 2), "d2h": round(d, 2)}), flush=True)
