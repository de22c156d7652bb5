"""Where the time of one pipelined exchange (kem_step_io) goes: the same 7-in / 4-out exchange
of 1e7 hh_ideal DOFs with parts of it switched off.

    python tools/exchange_probe.py [n_dofs]

Environment knobs of the runtime that matter here: KNPEMI_IO_CHUNKS, KNPEMI_IO_TAPER."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np  # noqa: E402
from knpemi_b200 import _cabi  # noqa: E402
from knpemi_b200.ducks import PointSpace  # noqa: E402
from knpemi_b200.odeSolver import MembraneModel  # noqa: E402
from workloads import SETUP, builtin, load_tables, synthetic_tables  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
name = "hh_ideal"
ode = builtin(name)
S, P, X, mask = synthetic_tables(name, n)
stim = {"stim_amplitude": SETUP[name]["stim"]}
loc = lambda x: x[0] < 20e-6      # noqa: E731


def pinned(src=None):
    a = _cabi.pinned_empty(n)
    a[:] = 0.0 if src is None else src
    return a


ins_all = {("parameter", k): pinned(P[:, ode.parameter_indices(k)]) for k in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")}
ins_all[("state", "V")] = pinned(S[:, 3])
outs_all = {("state", "V"): pinned(), **{("parameter", k): pinned() for k in ("I_ch_Na", "I_ch_K", "I_ch_Cl")}}


def run(label, n_sub, ins, outs, reps=8):
    m = MembraneModel(ode, None, 1, PointSpace(X), devices=[0], verbose=False, n_sub=n_sub, unread_inputs="discard")
    load_tables(m, S, P)
    for _ in range(3):
        m.step_exchange(SETUP[name]["dt"], ins, outs, stim, loc)
    t0 = time.perf_counter()
    acc = {"ms_total": 0.0, "ms_h2d": 0.0, "ms_kernel": 0.0, "ms_d2h": 0.0}
    for _ in range(reps):
        tm = m.step_exchange(SETUP[name]["dt"], ins, outs, stim, loc)
        for k in acc:
            acc[k] += tm[k] / reps
    wall = (time.perf_counter() - t0) * 1e3 / reps
    m.close()
    print(json.dumps({"case": label, "wall_ms": round(wall, 3), **{k: round(v, 3) for k, v in acc.items()}}), flush=True)


print(json.dumps({"n": n, "chunks": os.environ.get("KNPEMI_IO_CHUNKS", "16"), "taper": os.environ.get("KNPEMI_IO_TAPER", "1")}))
run("full: 7 in, 4 out, RK4 x 25", 25, ins_all, outs_all)
run("no kernel work: 7 in, 4 out, RK4 x 1", 1, ins_all, outs_all)
run("inputs only: 7 in, 0 out, RK4 x 25", 25, ins_all, {})
run("inputs only, no kernel work: RK4 x 1", 1, ins_all, {})
run("outputs only: 0 in, 4 out, RK4 x 25", 25, {}, outs_all)
run("outputs only, no kernel work: RK4 x 1", 1, {}, outs_all)
