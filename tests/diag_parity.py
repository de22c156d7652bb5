"""Per-column error report of the CUDA kernel against the oracle (test infrastructure: it uses
the oracle, so it lives under tests/).  Usage: python tests/diag_parity.py <model> <n> <steps>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from test_gpu_parity import run_pair
from workloads import builtin

name = sys.argv[1]; n = int(float(sys.argv[2])); steps = int(sys.argv[3])
for math in ("libm", "fast"):
    gS, gP, S, P = run_pair(name, n, steps, math=math)
    ode = builtin(name)
    print(f"== {name} math={math}")
    for c, (nm, _) in enumerate(ode.STATES):
        e = np.abs(gS[:, c] - S[:, c]); k = np.argmax(e / np.maximum(np.abs(S[:, c]), 1e-300))
        print(f"  state {nm:8s} max abs {e.max():.3e} max rel {np.max(e/np.maximum(np.abs(S[:,c]),1e-300)):.3e} (value {S[k,c]:.6e}) colmax {np.abs(S[:,c]).max():.3e}")
    for c, (nm, _) in enumerate(ode.PARAMETERS):
        e = np.abs(gP[:, c] - P[:, c])
        if e.max() == 0: continue
        k = np.argmax(e / np.maximum(np.abs(P[:, c]), 1e-300))
        print(f"  param {nm:8s} max abs {e.max():.3e} max rel {np.max(e/np.maximum(np.abs(P[:,c]),1e-300)):.3e} (value {P[k,c]:.6e}) colmax {np.abs(P[:,c]).max():.3e}")
