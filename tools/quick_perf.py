"""Quick kernel-only throughput probe (development tool, not the benchmark)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np
from knpemi_b200 import _cabi
from knpemi_b200.ducks import PointSpace
from knpemi_b200.models import BUILTIN
from knpemi_b200.odeSolver import MembraneModel
sys.path.insert(0, os.path.join(ROOT, "tests"))
from workloads import synthetic_tables, SETUP   # noqa

def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(BUILTIN)
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
    blocks = [int(b) for b in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
    scheme = sys.argv[4] if len(sys.argv) > 4 else "rk4"
    nvcc_flags = tuple(sys.argv[5].split()) if len(sys.argv) > 5 else ()
    tf, ms = _cabi.fp64_peak(0)
    print(f"fp64 DFMA peak: {tf:.2f} TFLOP/s ({ms:.3f} ms/launch); hbm copy {_cabi.hbm_copy_peak(0):.0f} GB/s")
    for name in names:
        ode = BUILTIN[name]
        S, P, X, mask = synthetic_tables(name, n, seed=20240611)
        for block in blocks:
            m = MembraneModel(ode, None, 1, PointSpace(X), devices=[0], verbose=False, block=block, scheme=scheme, nvcc_flags=nvcc_flags)
            for c in range(S.shape[1]):
                m.states[:, c] = S[:, c]
            for c in range(P.shape[1]):
                if np.all(P[:, c] == P[0, c]):
                    m.set_parameter_values({_pname(ode, c): (lambda x, v=P[0, c]: v)})
                else:
                    m._set_column(1, c, np.ascontiguousarray(P[:, c]))
            if os.environ.get("KNPEMI_NO_ACTIVITY_SORT"):
                m.set_activity_sort(False)
            dt = SETUP[name]["dt"]
            stim = {"stim_amplitude": SETUP[name]["stim"]}
            loc = lambda x: x[0] < 20e-6
            ts = []
            for k in range(8):
                m.step(dt, stim, loc, timed=True)
                ts.append(m.last_step_times["ms_kernel"])
            best, med = min(ts[2:]), float(np.median(ts[2:]))
            info = m.launch_info(block)
            if scheme == "dp45":
                a, r = m.step_stats()
                print(f"   dp45: {a / n / 8:.2f} accepted + {r / n / 8:.2f} rejected steps per DOF-step "
                      f"= {6 * (a + r) / n / 8 + 1:.1f} RHS evaluations (rk4 x 25: 101)")
            print(f"{name:13s} N={n:.0e} block={block or 'def'} regs={info['registers_per_thread']} "
                  f"blk/SM={info['blocks_per_sm']} kernel {med:.3f} ms (best {best:.3f}) -> "
                  f"{n / med * 1e3:.3e} DOF-steps/s")
            m.close()

def _pname(ode, c):
    # reverse lookup of a parameter name from its column
    for nm, _ in getattr(ode, "PARAMETERS"):
        if ode.parameter_indices(nm) == c:
            return nm
    raise KeyError(c)

if __name__ == "__main__":
    main()
