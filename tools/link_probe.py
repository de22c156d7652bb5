"""Host-link ceilings of one B200 (pinned copies, both directions, several copy sizes).

    python tools/link_probe.py [dev]

Prints one JSON object per copy size: what `bench.py` divides the end-to-end exchange by."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "knp-emi-fenics-x_b200"))
from knpemi_b200 import _cabi  # noqa: E402

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
for mb in (1, 5, 16, 80):
    r = _cabi.link_ceiling(dev, mb << 20, reps=max(4, 400 // mb if mb > 4 else 64))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}))
# asymmetric loads: what the exchange of one PDE step looks like (5 columns in, 3 or 4 out)
for rin, rout in ((40, 24), (40, 32), (56, 32)):
    h, d = _cabi.link_probe(dev, 5 << 20, rin, rout)
    print(json.dumps({"copy_bytes": 5 << 20, "reps_h2d": rin, "reps_d2h": rout, "h2d": round(h, 2), "d2h": round(d, 2)}))
