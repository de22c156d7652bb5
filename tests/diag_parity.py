"""Strict per-column parity report of the CUDA kernel against the oracle (test infrastructure:
it uses the oracle, so it lives under tests/).

    python tests/diag_parity.py [model|all] [n] [steps] [--json out.json]

For every state and output column it reports, for the product build (math="fast") and the
triage build (math="libm": CUDA libm + IEEE division, differs from the oracle only by FMA
contraction and CUDA-vs-glibc libm):

  strict_max_rel    max |got - want| / |want| over the entries with want != 0 (no floor)
  n_above_tol       entries whose strict relative error exceeds 1e-10
  max_abs           max |got - want|
  abs_over_colmax   max_abs / max |want| of the column
  min_floor_frac    the smallest f for which max |got - want| / max(|want|, f * colmax) < 1e-10
                    (0 if the strict error already passes)

north_star's bar is 1e-10 relative "on all states and I_ch".  States pass it strictly wherever
they are not crossing zero.  A channel current is g*(V - E) + pump terms: its error is the
state error times the conductance, about 1e-13 of the column's scale, at every DOF -- also at
the DOFs where the terms cancel and the current itself is 1e-5 of that scale.  Those entries
are ill-conditioned in any arithmetic: the libm build shows the same numbers.  The tests
therefore measure currents against max(|entry|, 1e-3 * column max) and states against
max(|entry|, 1e-6 * column max); this report is what justifies the two floors.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np  # noqa: E402

TOL = 1e-10
STATE_FLOOR, CURRENT_FLOOR = 1e-6, 1e-3


def column_report(got, want):
    err = np.abs(got - want)
    colmax = float(np.max(np.abs(want)))
    nz = want != 0
    strict = err[nz] / np.abs(want[nz]) if nz.any() else np.zeros(0)
    rep = {"strict_max_rel": float(strict.max()) if strict.size else 0.0,
           "n_above_tol": int(np.sum(strict > TOL)), "n": int(want.size),
           "max_abs": float(err.max()), "colmax": colmax,
           "abs_over_colmax": float(err.max() / colmax) if colmax > 0 else 0.0}
    # smallest floor fraction that brings every entry under TOL: entry i needs
    # max(|want_i|, f colmax) > err_i / TOL, i.e. f > err_i / (TOL colmax) where |want_i| is too small
    need = err / TOL
    short = need > np.abs(want)
    rep["min_floor_frac"] = float(np.max(need[short]) / colmax) if short.any() and colmax > 0 else 0.0
    return rep


def parity_report(name, n=20000, steps=10):
    from test_gpu_parity import run_pair
    from workloads import builtin
    ode = builtin(name)
    out = {"model": name, "dofs": n, "pde_steps": steps, "tolerance": TOL, "builds": {}}
    for math in ("fast", "libm"):
        gS, gP, S, P = run_pair(name, n, steps, math=math)
        cols = {}
        for c, (nm, _) in enumerate(ode.STATES):
            cols["state " + nm] = column_report(gS[:, c], S[:, c])
        for c, (nm, _) in enumerate(ode.PARAMETERS):
            if np.any(gP[:, c] != P[:, c]) or nm.startswith("I_ch"):
                cols["parameter " + nm] = column_report(gP[:, c], P[:, c])
        out["builds"][math] = cols
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "all"
    n = int(float(args[1])) if len(args) > 1 else 20000
    steps = int(args[2]) if len(args) > 2 else 10
    from test_gpu_parity import MODELS
    names = MODELS if name == "all" else (name,)
    reports = [parity_report(m, n, steps) for m in names]
    for r in reports:
        for math, cols in r["builds"].items():
            print(f"== {r['model']} math={math}  ({r['dofs']} DOFs x {r['pde_steps']} PDE steps)")
            for col, e in cols.items():
                print(f"  {col:22s} strict max rel {e['strict_max_rel']:.2e}  above 1e-10: {e['n_above_tol']:6d}/{e['n']}"
                      f"  max abs {e['max_abs']:.2e} = {e['abs_over_colmax']:.1e} of column max"
                      f"  needs floor {e['min_floor_frac']:.1e}")
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        with open(path, "w") as f:
            json.dump(reports, f, indent=1)


if __name__ == "__main__":
    main()
