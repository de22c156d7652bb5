"""What codegen/fuse_exp.py does to each builtin model: in-loop operation counts and groups."""
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "knp-emi-fenics-x_b200"))
from knpemi_b200 import models                       # noqa: E402
from knpemi_b200.codegen import emit, parse          # noqa: E402

for name in models.BUILTIN:
    mod = importlib.import_module(f"knpemi_b200.models.{name}")
    pm = parse.parse_model_source(open(mod.__file__).read(), mod.__file__)
    ns, np_ = len(mod.init_state_values()), len(mod.init_parameter_values())
    a = emit.emit_model(pm, name, ns, np_, emit.EmitOptions(fuse_exp=False))
    b = emit.emit_model(pm, name, ns, np_)
    print(name, a.stats["deriv"], "->", b.stats["deriv"])
    for r in b.stats["fused_exp"]:
        print("   ", r)
