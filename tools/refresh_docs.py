"""Rewrite docs/generated_<model>.cu from the generator (tests/test_codegen.py checks they are current)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "knp-emi-fenics-x_b200")]
from knpemi_b200 import codegen  # noqa: E402
from knpemi_b200.models import BUILTIN  # noqa: E402

for name in ("hh_ideal", "calibration"):
    with open(os.path.join(ROOT, "docs", f"generated_{name}.cu"), "w") as f:
        f.write(codegen.generate(BUILTIN[name]).source)
    print("wrote docs/generated_%s.cu" % name)
