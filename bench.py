#!/usr/bin/env python
"""Benchmark of the membrane-ODE stage: membrane DOF-steps/s (fp64) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A *step* is one PDE step of the membrane stage over every DOF of the workload
(MembraneModel.step_lsoda: sticky stimulus, RK4 x n_sub, current epilogue).  The
default workload is BASELINE.json configs[2], the configuration the metric
("... at 1/2/4/8 B200") is quoted on: idealized HH (`hh_ideal`) with the six
interface concentrations as per-DOF inputs, 10^7 membrane DOFs per GPU (weak
scaling: DOFs are independent, ranges are disjoint, no collective on the data
path).  Its working set (1.1 GB per GPU) is larger than the 126 MB L2, so no L2
flush is needed between iterations.

One JSON line on stdout (rank 0).  `value` is measured with inputs resident in
HBM and CUDA events on the launching stream (`sustained`: the same loop for at
least a second); `e2e` goes through the public MembraneModel API
(`step_exchange`) with pinned HOST buffers, host<->device copies inside the
timed region -- `e2e.link` is the same copies with nothing else (the host-link
floor, measured in the same run, all ranks at once) and `e2e.dropin` the
reference's unmodified 7-setter + step + 4-getter call sequence; `roofline`
states the FP64-pipe issue-slot utilisation of the fused kernel against a DFMA
peak measured in the same run (MEASURED_PEAKS.json has no fp64 entry) plus the
HBM view; `parity_sample` steps 10^5 random DOFs of the workload on the GPU
and in the oracle; `cpu_baseline` is the oracle port of the reference's
stepping on the host cores (rank 0, N = 1 only).  `--workload tissue_1e8` is
BASELINE.json configs[4]: two membrane models, 10^8 DOFs, strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

_emit = print
METRIC = "membrane DOF-steps/sec (fp64)"
UNIT = "DOF-steps/s"
N_SUB = 25

WORKLOADS = {
    # name: (model, DOFs per GPU, BASELINE.json config it realises)
    "hh_ideal_1e7": ("hh_ideal", 10_000_000, "configs[2]: 3D idealized neuron, HH + Na/K/Cl interface "
                                             "concentrations, 10^7 membrane DOFs per GPU"),
    "hh_test_1e6": ("hh_test", 1_000_000, "configs[1]: tests/mm_test_ode.py HH system, 10^6 DOFs"),
    "calibration_1e7": ("calibration", 10_000_000, "configs[3]: calibration ODE system, 10^7 DOFs"),
    "hh_tissue_1e7": ("hh_tissue", 10_000_000, "configs[4] neuron part: tissue HH, 10^7 DOFs per GPU"),
    "glial_tissue_1e7": ("glial_tissue", 10_000_000, "configs[4] glial part: mm_glial, 10^7 DOFs per GPU"),
    # config #5: two membrane models (neurons tag 1, glia tag 2: run_stim_duration.py:171-181),
    # 10^8 DOFs in total, strong-sharded over the GPUs of the run
    "tissue_1e8": ([("hh_tissue", 50_000_000), ("glial_tissue", 50_000_000)],
                   "configs[4]: EMIx-scale tissue membrane, HH (tag 1) + glial (tag 2), 10^8 DOFs in total, "
                   "contiguous ranges of both models over the GPUs"),
}

# algorithmic HBM bytes per DOF-step: columns read + columns written, 8 B each
ALGO_BYTES = {"hh_ideal": 144, "hh_tissue": 144, "hh_test": 96, "glial_tissue": 88, "glial_bench": 88,
              "calibration": 224}
IO_COLUMNS = {   # what crosses the host link per PDE step in drop-in mode (utils.py:217-233, run_2D.py:105-109)
    "hh_ideal": (["K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i"], True, ["I_ch_Na", "I_ch_K", "I_ch_Cl"]),
    "hh_tissue": (["K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i"], True, ["I_ch_Na", "I_ch_K", "I_ch_Cl"]),
    "glial_tissue": (["K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i"], True, ["I_ch_Na", "I_ch_K", "I_ch_Cl"]),
    "hh_test": ([], True, ["I_ch_Na", "I_ch_K", "I_ch_Cl"]),
    "calibration": ([], False, []),
}


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """`nvidia-smi` SM clocks and throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc, self.thread = gpu_index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()
        # nvidia-smi needs 0.1-0.5 s before its first row; a timed region of 30 steps (0.15 s) that
        # starts earlier can end without a single sample
        self._wait(lambda: bool(self.rows), 3.0)

    def _wait(self, done, timeout):
        t_end = time.perf_counter() + timeout
        while not done() and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self._wait(lambda: bool(self.rows) and self.rows[-1][0] > t1, 0.5)   # the sample that closes the region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = list(self.rows)
        inside = [r for r in rows if t0 <= r[0] <= t1 + 0.06]
        bracketing = not inside
        if bracketing:      # region shorter than one sampling period: the samples either side of it
            inside = [r for r in rows if r[0] < t0][-1:] + [r for r in rows if r[0] > t1][:1]
        for ts, line in inside:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons), "samples": len(sm),
                **({"bracketing_samples": True} if bracketing and sm else {})}


class Dist:
    """torch.distributed plumbing for N > 1 (barrier + max over ranks); no-op for N = 1."""

    def __init__(self, init=True):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        self.backend = None
        if self.world > 1 and init:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            backend = os.environ.get("KNPEMI_BENCH_BACKEND", "nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {"device_id": torch.device("cuda", self.local_rank)} if backend == "nccl" else {}
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world, **kw)
            self.backend = backend

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max(self, x: float) -> float:
        if self.world == 1:
            return x
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = self.torch.tensor([x], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if self.world == 1:
            return x
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = self.torch.tensor([x], dtype=self.torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, x: float) -> list:
        """Value of every rank (sum-reduction of one-hot vectors)."""
        if self.world == 1:
            return [x]
        dev = "cuda" if self.backend == "nccl" else "cpu"
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=dev)
        t[self.rank] = x
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def close(self):
        if self.world > 1 and self.torch is not None:
            self.dist.destroy_process_group()


def fp64_slots(model_name: str, n_sub: int):
    """FP64-pipe thread-instructions per DOF-step of the model's fused kernel.

    profiles/fp64_slots.json holds, per model, the static SASS count (tools/sass_stats.py:
    loop body x 4 n_sub + the rest once) and, where captured, ncu's dynamic count
    (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum / N)."""
    path = os.path.join(ROOT, "profiles", "fp64_slots.json")
    try:
        with open(path) as f:
            table = json.load(f)
        e = table[model_name]
    except (OSError, KeyError, ValueError):
        return None, None
    if e.get("ncu_per_dof_step") and e.get("ncu_n_sub") == n_sub:
        return float(e["ncu_per_dof_step"]), "ncu"
    iters = 4 // int(e.get("stages_per_loop_iteration", 1))
    return float(iters * n_sub * e["loop_fp64"] + e["once_fp64"]), "sass-static"


def host_threads() -> int:
    """Every core this process may use -- not OMP_NUM_THREADS, which torchrun sets to 1."""
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, 1)


def l2_note(parts) -> str:
    """Timing rule: inputs larger than the 126 MB L2, or say that they are not."""
    mb = sum(ALGO_BYTES[name] * n for name, n in parts) / 1e6
    if mb > 126:
        return f"inputs {mb:.0f} MB per GPU > 126 MB L2, no flush needed"
    return (f"inputs {mb:.0f} MB per GPU < 126 MB L2 and NOT flushed: informational workload, not a "
            "contract line (compute-bound kernel)")


def workload_parts(workload: str):
    """[(model name, DOFs per GPU or in total)], description, strong?"""
    spec = WORKLOADS[workload]
    if isinstance(spec[0], (list, tuple)):
        return list(spec[0]), spec[1], True
    return [(spec[0], spec[1])], spec[2], False


def cpu_port_rate(workload: str, steps: int, warmup: int, budget_s: float, seed: int = 20240611):
    """The oracle port of the reference stepping (oracle/knpemi_oracle.c: the reference's
    right-hand sides bit for bit, the row loop of odeSolver.py:107-122, scheme O1) on every
    host thread, on a bounded sample of the workload split over its membrane models like the
    workload itself.  One procedure for the reference arm and the in-arm `cpu_baseline`:
    (DOF-steps/s, threads, sample text, seconds per step)."""
    from oracle import cpu_oracle
    from workloads import SETUP, synthetic_tables
    threads = host_threads()
    parts, _, _ = workload_parts(workload)
    total = float(sum(n for _, n in parts))
    probe = []
    for name, n in parts:                       # cold probe: sizes the sample, not reported
        c = max(int(4000 * threads * n / total), 1000)
        S, P, X, mask = synthetic_tables(name, c, seed)
        t0 = time.perf_counter()
        cpu_oracle.step(name, S, P, 0.0, SETUP[name]["dt"], N_SUB, threads)
        probe.append((c, time.perf_counter() - t0))
    rate = sum(c for c, _ in probe) / max(sum(t for _, t in probe), 1e-9)
    n_all = int(min(max(rate * budget_s / max(steps + warmup, 1), 4000 * threads), 4_000_000))
    tabs = []
    for name, n in parts:
        k = max(int(n_all * n / total), 1000)
        S, P, X, mask = synthetic_tables(name, k, seed)
        P[mask, _stim_col(name)] = SETUP[name]["stim"]
        tabs.append((name, k, S, P))
    t = {name: 0.0 for name, _ in parts}

    def one_step():
        for name, k, S, P in tabs:
            bad = cpu_oracle.step(name, S, P, t[name], SETUP[name]["dt"], N_SUB, threads)
            assert bad == 0
            t[name] += SETUP[name]["dt"]
    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    el = time.perf_counter() - t0
    n_used = sum(k for _, k, _, _ in tabs)
    sample = (" + ".join(f"{k} DOFs of {name}" for name, k, _, _ in tabs)
              + f" per step (RK4 x {N_SUB}), {steps} steps after {warmup} warm-up, {el:.1f} s")
    return n_used * steps / el, threads, sample, el / steps


def run_cpu_lsoda(model_name: str, n: int, seed: int):
    """B2 of BASELINE.md: the reference's own semantics -- serial Python loop over rows, one
    cold-started LSODA solve per row at rtol 1e-8 / atol 1e-10 (odeSolver.py:107-122) -- with
    scipy's LSODA standing in for the absent numbalsoda.  One core, small N; the Python
    callback per RHS evaluation makes this 10-100x slower than numbalsoda would be."""
    from ducks_for_tests import Space
    from oracle.membrane_oracle import OracleMembraneModel
    from workloads import SETUP, builtin, synthetic_tables
    S, P, X, mask = synthetic_tables(model_name, n, seed)
    m = OracleMembraneModel(builtin(model_name), None, 1, Space(X), oracle_name=model_name)
    m.states[:] = S
    m.parameters[:] = P
    cfg = SETUP[model_name]
    t0 = time.perf_counter()
    m.step_lsoda_scipy(cfg["dt"], {"stim_amplitude": cfg["stim"]}, lambda x: x[0] < 20e-6)
    el = time.perf_counter() - t0
    return {"value": n / el, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} DOFs x 1 PDE step of {model_name}, scipy LSODA per row, {el:.1f} s",
            "caveat": "Python callback per RHS evaluation; numbalsoda itself would be 10-100x faster"}


def _stim_col(model_name):
    from workloads import builtin
    return builtin(model_name).parameter_indices("stim_amplitude")


# ------------------------------------------------------------------ reference arm
def run_reference(args, dist: Dist):
    """The reference's CPU stepping on the host cores, same metric and config.

    The reference (Python + numba cfuncs + numbalsoda) cannot be installed here:
    numbalsoda and dolfinx are absent from the offline wheelhouse.  The arm therefore times
    the oracle port (oracle/knpemi_oracle.c: the reference's right-hand sides bit-for-bit,
    the row loop of odeSolver.py:107-122, scheme O1) with every host thread OpenMP gives it.
    Each step is a bounded sample of the workload (rank 0 only)."""
    if dist.rank != 0:
        return
    from workloads import SETUP
    parts, cfg_text, strong = workload_parts(args.workload)
    budget_s = float(os.environ.get("KNPEMI_BENCH_REF_BUDGET_S", "150"))      # whole run within ~2.5 min
    value, threads, sample, s_per_step = cpu_port_rate(args.workload, args.steps, max(args.warmup, 1), budget_s)
    head = parts[0][0]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload,
                   "membrane_model": head if not strong else [name for name, _ in parts],
                   "baseline_config": cfg_text,
                   "dofs_per_gpu": parts[0][1] if not strong else None,
                   "scheme": "rk4", "n_sub": N_SUB, "dt": SETUP[head]["dt"], "l2": l2_note(parts)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU stepping; timed: oracle port of odeSolver.py:107-122 with the reference's "
                "RHS, OpenMP over rows, all host threads (numbalsoda/dolfinx not installable offline)",
    }
    _emit(json.dumps(line))


# ------------------------------------------------------------------------ GPU arm
class Part:
    """One MembraneModel of the workload (config #5 has two: neurons tag 1, glia tag 2)."""

    def __init__(self, model_name, n, dev, seed, args, **model_kw):
        from knpemi_b200.ducks import PointSpace
        from knpemi_b200.odeSolver import MembraneModel
        from workloads import SETUP, builtin, load_tables, synthetic_tables
        self.name, self.n, self.cfg, self.ode = model_name, n, SETUP[model_name], builtin(model_name)
        self.S, self.P, self.X, self.mask = synthetic_tables(model_name, n, seed=seed)
        self.model = MembraneModel(self.ode, None, 1, PointSpace(self.X), devices=[dev], verbose=False,
                                   n_sub=N_SUB, block=args.block, scheme=args.scheme, **model_kw)
        load_tables(self.model, self.S, self.P)
        self.stim = {"stim_amplitude": self.cfg["stim"]}
        self.locator = lambda x: x[0] < 20e-6            # noqa: E731  (run_2D.py:264)
        self.dt = self.cfg["dt"]
        self.keep = []

    def step_async(self):
        self.model.step_async(self.dt, self.stim, self.locator)

    # -- host buffers of one PDE<->ODE exchange (utils.py:217-233, run_2D.py:105-109)
    def make_exchange_buffers(self, pinned=True):
        from knpemi_b200 import _cabi
        in_names, v_io, out_names = IO_COLUMNS[self.name]

        def buf(src=None):
            a = _cabi.pinned_empty(self.n) if pinned else np.empty(self.n)
            a[:] = 0.0 if src is None else src
            self.keep.append(a)
            return a

        self.ins = {("parameter", k): buf(self.P[:, self.ode.parameter_indices(k)]) for k in in_names}
        self.outs = {("parameter", k): buf() for k in out_names}
        self.v_io = v_io
        if v_io:
            self.ins[("state", "V")] = buf(np.asarray(self.model.states[:, self.ode.state_indices("V")]))
            self.outs[("state", "V")] = buf()

    def exchange(self):
        tm = self.model.step_exchange(self.dt, self.ins, self.outs, self.stim, self.locator)
        if self.v_io:                                 # the PDE side would hand phi_M back
            self.ins[("state", "V")], self.outs[("state", "V")] = self.outs[("state", "V")], self.ins[("state", "V")]
        return tm

    def link_bytes(self):
        """(bytes offered in, bytes that crossed host->device, bytes copied device->host, bytes the
        host filled itself) per exchange, from where the library says the columns live."""
        m = self.model
        kept = [k for (what, k) in self.ins if what == "parameter" and m.column_location(what, k) in ("host", "discarded")]
        literal = [k for (what, k) in self.outs if what == "parameter" and k in self.literal_outputs()]
        return (8 * self.n * len(self.ins), 8 * self.n * (len(self.ins) - len(kept)),
                8 * self.n * (len(self.outs) - len(literal)), 8 * self.n * len(literal), kept, literal)

    def literal_outputs(self):
        em = self.model._emitted
        src = em.source
        import re
        m = re.search(r"const int CONST_OUT_COLS\[\] = \{([^}]*)\};", src)
        n_const = int(re.search(r"USED_COLS, \d+, (\d+), CONST_OUT_COLS", src).group(1))
        cols = [int(x) for x in m.group(1).split(",")][:n_const]
        names = []
        for (what, k) in self.outs:
            if what == "parameter" and self.ode.parameter_indices(k) in cols:
                names.append(k)
        return names

    def parity_sample(self, rows, steps=2):
        """`rows` random DOFs of this part stepped by the CUDA path and by the CPU oracle from the
        same tables: max relative error (the rule of tests/test_gpu_parity.py).  bench.py may call
        the oracle only as a checker; this is that."""
        from knpemi_b200.ducks import PointSpace
        from knpemi_b200.odeSolver import MembraneModel
        from oracle import cpu_oracle
        from workloads import load_tables
        rng = np.random.default_rng(5)
        idx = np.sort(rng.choice(self.n, size=min(rows, self.n), replace=False))
        S, P, X = self.S[idx].copy(), self.P[idx].copy(), self.X[idx]
        mask = self.mask[idx]
        m = MembraneModel(self.ode, None, 1, PointSpace(X), devices=self.model.devices, verbose=False, n_sub=N_SUB)
        load_tables(m, S, P)
        t = 0.0
        for _ in range(steps):
            m.step_lsoda(self.dt, self.stim, self.locator)
            P[mask, self.ode.parameter_indices("stim_amplitude")] = self.cfg["stim"]
            assert cpu_oracle.step(self.name, S, P, t, self.dt, N_SUB) == 0
            t += self.dt
        got_s = np.asarray(m.states)
        out_cols = m.output_columns
        got_c = np.asarray(m.parameters)[:, out_cols] if out_cols else np.zeros((len(idx), 0))
        m.close()

        def rel(a, b, floor):
            if b.size == 0:
                return 0.0
            scale = np.maximum(np.abs(b), floor * np.max(np.abs(b), axis=0, keepdims=True) + 1e-300)
            return float(np.max(np.abs(a - b) / scale))
        # the floors of tests/test_gpu_parity.py (states 1e-6, currents 3e-4 of the column maximum;
        # profiles/r2_parity_strict.md says why a current needs one)
        return {"rows": int(len(idx)), "pde_steps": steps, "max_rel_err_states": rel(got_s, S, 1e-6),
                "max_rel_err_currents": rel(got_c, P[:, out_cols], 3e-4), "tolerance": 1e-10,
                "floor": "states 1e-6, currents 3e-4 x column max"}


def measure_link(dev, dist, copy_bytes, reps_h2d, reps_d2h, cycles=20):
    """Host-link ceiling for THIS exchange: the same pinned copies one PDE step makes -- the same
    number, size and direction -- with nothing else: no kernel, no dependency between them,
    streaming through `span` bytes of host memory per direction (far above the last-level cache,
    like the caller's arrays).  `cycles` exchanges' worth of copies are issued back to back and
    the time is divided by `cycles`: a single exchange's worth is a burst the host absorbs at up
    to twice its sustained rate when eight GPUs share it (profiles/r2_exchange.md), and the
    bench's exchanges run back to back.  Under torchrun every rank of the box does it at the same
    time.  Returns per-direction rates and the time per exchange of the slowest rank: no
    exchange of these bytes over this host link, sustained, can be faster."""
    from knpemi_b200 import _cabi
    span = 480 << 20
    out = {"copy_bytes": copy_bytes, "copies_h2d_per_exchange": reps_h2d, "copies_d2h_per_exchange": reps_d2h,
           "exchanges_back_to_back": cycles, "host_span_bytes_per_direction": span}
    if dist.rank == 0:
        r = _cabi.link_ceiling(dev, 5 << 20, 96, span)
        out["one_gpu_each_direction"] = {k: round(v, 2) for k, v in r.items() if isinstance(v, float)}
    best = None
    for _ in range(2):                       # the better of two: the floor is a lower bound on time
        dist.barrier()
        h, d = _cabi.link_probe(dev, copy_bytes, reps_h2d * cycles, reps_d2h * cycles, span)
        ms = max(copy_bytes * reps_h2d / (h * 1e6) if reps_h2d else 0.0,
                 copy_bytes * reps_d2h / (d * 1e6) if reps_d2h else 0.0)
        ms_all = dist.max(ms)
        if best is None or ms_all < best[0]:
            best = (ms_all, dist.gather(h), dist.gather(d), dist.gather(ms))
    out["same_copies_all_ranks_concurrent"] = {
        "h2d_gbs_per_rank": [round(v, 1) for v in best[1]], "d2h_gbs_per_rank": [round(v, 1) for v in best[2]],
        "ms_per_exchange_per_rank": [round(v, 3) for v in best[3]], "ms": best[0],
        "h2d_gbs_sum": round(sum(best[1]), 1), "d2h_gbs_sum": round(sum(best[2]), 1)}
    return out


def run_gpu(args, dist: Dist):
    from knpemi_b200 import _cabi
    from knpemi_b200.ducks import ArrayFunction

    part_specs, cfg_text, strong = workload_parts(args.workload)
    dev = dist.local_rank
    if _cabi.device_count() <= dev:
        raise SystemExit("bench.py: no CUDA device for this rank -- the product has no CPU path")
    # several ranks on a multi-socket host: keep this rank's pinned buffers next to its GPU
    cpus = []
    if dist.world > 1 and not os.environ.get("KNPEMI_NO_AFFINITY"):
        from knpemi_b200.affinity import bind_to_device
        cpus = bind_to_device(dev)

    if strong:
        # config #5: fixed total size, contiguous DOF ranges of every membrane model over the ranks
        parts = []
        for k, (model_name, n_total) in enumerate(part_specs):
            from knpemi_b200.sharding import rank_range     # the rule kem_create applies to its devices
            n_total = int(args.dofs) if args.dofs else n_total
            lo, hi = rank_range(n_total, dist.rank, dist.world)
            parts.append(Part(model_name, hi - lo, dev, 20240611 + 97 * k + dist.rank, args,
                              unread_inputs=args.unread_inputs))
    else:
        model_name, n = part_specs[0]
        if args.dofs:
            n = int(args.dofs)
        parts = [Part(model_name, n, dev, 20240611 + dist.rank, args, unread_inputs=args.unread_inputs)]
    head = max(parts, key=lambda p: p.n * (fp64_slots(p.name, N_SUB)[0] or 1.0))   # dominant kernel
    n_rank = sum(p.n for p in parts)

    # ------------------------------------------------ resident: inputs already in HBM
    def resident(steps):
        for p in parts:
            p.model.timer_begin()
        for _ in range(steps):
            for p in parts:
                p.step_async()
        return max(p.model.timer_end() for p in parts)

    for _ in range(max(args.warmup, 3)):
        for p in parts:
            p.step_async()
    for p in parts:
        p.model.synchronize()
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.12)
    dist.barrier()
    launches0 = sum(p.model.launch_count() for p in parts)
    wall0 = time.perf_counter()
    ms = resident(args.steps)
    for p in parts:
        p.model.synchronize()
    wall1 = time.perf_counter()
    dist.barrier()
    launches = sum(p.model.launch_count() for p in parts) - launches0
    clocks = sampler.stop(wall0, wall1)
    ms_max = dist.max(ms)
    total_dofs = dist.sum(float(n_rank))
    value = total_dofs * args.steps / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps

    # the same loop for at least a second: what the rate is once clocks and power have settled
    sustained = None
    if args.sustain_seconds > 0:
        k_sus = max(int(args.sustain_seconds * 1e3 / max(ms / args.steps, 1e-3)) + 1, args.steps)
        s2 = ClockSampler(dev)
        s2.start()
        time.sleep(0.06)
        dist.barrier()
        w0 = time.perf_counter()
        ms_sus = resident(k_sus)
        for p in parts:
            p.model.synchronize()
        w1 = time.perf_counter()
        ms_sus = dist.max(ms_sus)
        sustained = {"steps": k_sus, "seconds": ms_sus * 1e-3, "value": total_dofs * k_sus / (ms_sus * 1e-3),
                     "unit": UNIT, "clocks": s2.stop(w0, w1)}

    # kernel of the dominant part alone (roofline numerator's duration; equals ms/steps when
    # the workload has one part)
    if len(parts) > 1:
        head.model.timer_begin()
        for _ in range(args.steps):
            head.step_async()
        kernel_ms = head.model.timer_end() / args.steps
    else:
        kernel_ms = ms / args.steps

    dp45_steps = None
    if args.scheme == "dp45":
        head.model.step_stats()                                   # discard warm-up counts
        for _ in range(3):
            head.step_async()
        acc, rej = head.model.step_stats()
        dp45_steps = {"accepted_per_dof_step": acc / (3.0 * head.n), "rejected_per_dof_step": rej / (3.0 * head.n),
                      "rhs_evals_per_dof_step": 6.0 * (acc + rej) / (3.0 * head.n) + 1.0,
                      "rtol": head.model.rtol, "atol": head.model.atol}

    # ------------------------------------------------ end to end: host buffers through the API
    for p in parts:
        p.make_exchange_buffers(pinned=True)
    e2e_steps = max(min(args.steps, 20), 3)
    for _ in range(3):
        for p in parts:
            p.exchange()
    dist.barrier()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for p in parts:
            dev_ms += p.exchange()["ms_total"]
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    offered = h2d = d2h = filled = 0
    kept_names, literal_names = [], []
    for p in parts:
        o, i, d, f, kept, lit = p.link_bytes()
        offered, h2d, d2h, filled = offered + o, h2d + i, d2h + d, filled + f
        kept_names += [f"{p.name}.{k}" for k in kept]
        literal_names += [f"{p.name}.{k}" for k in lit]
    dist.barrier()
    e2e_ms_max = dist.max(e2e_wall_ms)
    e2e_per_rank = [round(v / e2e_steps, 3) for v in dist.gather(e2e_wall_ms)]
    e2e_dev_per_rank = [round(v / e2e_steps, 3) for v in dist.gather(dev_ms)]
    e2e_value = total_dofs * e2e_steps / (e2e_ms_max * 1e-3)
    last = dict(head.model.last_step_times)
    h2d_all, d2h_all = dist.sum(float(h2d)), dist.sum(float(d2h))

    # ------------------------------------------------ host link ceiling (denominator of e2e)
    # the copies of one exchange of the largest part, as the pipeline issues them: 16 chunks per column
    link, link_frac = None, None
    if not args.no_link_probe and head.n > 0 and h2d + d2h > 0:
        cols_in, cols_out = h2d // (8 * max(n_rank, 1)), d2h // (8 * max(n_rank, 1))
        chunk_bytes = max(8 * ((n_rank + 15) // 16), 1 << 16)
        link = measure_link(dev, dist, chunk_bytes, int(16 * cols_in), int(16 * cols_out))
        floor_ms = link["same_copies_all_ranks_concurrent"]["ms"]
        link_frac = {"floor_ms_per_step": floor_ms, "measured_ms_per_step": e2e_ms_max / e2e_steps,
                     "frac": floor_ms / (e2e_ms_max / e2e_steps),
                     "h2d_gbs_achieved": h2d_all / (e2e_ms_max / e2e_steps * 1e-3) / 1e9,
                     "d2h_gbs_achieved": d2h_all / (e2e_ms_max / e2e_steps * 1e-3) / 1e9,
                     "definition": "floor = time per exchange the same pinned copies (number, size, direction) take "
                                   "with no kernel and no dependencies, 20 exchanges' worth back to back, all "
                                   "ranks at once, slowest rank, measured in this run; frac = floor / measured "
                                   "exchange"}

    # ------------------------------------------------ the reference's own call sequence, unmodified:
    # 7 setter calls + step_lsoda + 4 getter calls per PDE step on ordinary NumPy arrays the
    # caller keeps (utils.py:227-233, run_2D.py:98-109) -- what a user sees without touching
    # solve_odes.  Arrays that come back every step are page-locked by the library on their
    # second sighting; `exchange="deferred"` is the one-keyword opt-in that pipelines the copies.
    dropin = {}
    p0 = parts[0]
    in_names, v_io, out_names = IO_COLUMNS[p0.name]
    if v_io and in_names and len(parts) == 1 and dist.world == 1 and not args.no_dropin:      # like cpu_baseline: N = 1 only
        from knpemi_b200.ducks import PointSpace
        from knpemi_b200.odeSolver import MembraneModel
        from workloads import load_tables
        n_warm, n_timed = 3, 5
        for mode in ("immediate", "deferred", "immediate_fresh_inputs"):
            m = MembraneModel(p0.ode, None, 1, PointSpace(p0.X), devices=[dev], verbose=False, n_sub=N_SUB,
                              exchange=mode.split("_")[0], unread_inputs=args.unread_inputs)
            load_tables(m, p0.S, p0.P)
            fresh = mode.endswith("fresh_inputs")
            # persistent caller arrays, or -- like the reference, whose interpolate_to_membrane makes
            # two new Functions per call (utils.py:190-191) -- a new input array at every step
            # (created before the clock starts: their allocation is the PDE side's)
            sets = [{k: ArrayFunction(p0.P[:, p0.ode.parameter_indices(k)].copy()) for k in in_names}
                    for _ in range(n_warm + n_timed if fresh else 1)]
            u_phi = ArrayFunction(p0.S[:, p0.ode.state_indices("V")].copy())
            u_out = {k: ArrayFunction(p0.n) for k in out_names}

            def one_pde_step(u_in):
                for k, u in u_in.items():
                    m.set_parameter(k, u)
                m.set_membrane_potential(u_phi)
                m.step_lsoda(p0.dt, p0.stim, p0.locator)
                m.get_membrane_potential(u_phi)
                for k, u in u_out.items():
                    m.get_parameter(k, u)

            for i in range(n_warm):       # the second step registers the arrays, the third runs on them
                one_pde_step(sets[i % len(sets)])
            dist.barrier()
            t0 = time.perf_counter()
            for i in range(n_timed):
                one_pde_step(sets[(n_warm + i) % len(sets)])
            call_ms = dist.max((time.perf_counter() - t0) * 1e3) / n_timed
            m.close()
            del sets
            dropin[mode] = {"value": total_dofs / (call_ms * 1e-3), "unit": UNIT, "ms_per_step_wall": call_ms,
                            "fraction_of_step_exchange": (e2e_ms_max / e2e_steps) / call_ms}
        dropin["api"] = ("set_parameter x6 + set_membrane_potential + step_lsoda + get_membrane_potential + "
                         "get_parameter x3 on ordinary NumPy arrays; 'immediate' = defaults with arrays the caller keeps "
                         "(page-locked by the library on their second sighting), 'deferred' = the same with "
                         "MembraneModel(exchange='deferred'), 'immediate_fresh_inputs' = defaults with a new array "
                         "for each of the six traces at every step, as the reference's interpolate_to_membrane "
                         "produces them (staged through pinned buffers; only phi_M and the outputs come back)")

    # ------------------------------------------------ roofline of the fused kernel (dominant part)
    peak_tf, _ = _cabi.fp64_peak(dev)
    slots, slots_src = fp64_slots(head.name, N_SUB)
    if dp45_steps and slots:
        # O3 spends a data-dependent number of RHS evaluations: scale the per-RHS count of the
        # O1 kernel (same generated right-hand side) by the measured evaluations per DOF-step
        slots = slots / (4 * N_SUB + 1) * dp45_steps["rhs_evals_per_dof_step"]
        slots_src += " x measured RHS evaluations of scheme O3 (excludes controller arithmetic)"
    roofline = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tf, "kernel": f"kem_step_kernel<{head.name}>",
                "peak_source": "measured in this run: kem_fp64_peak (8 independent DFMA chains/thread, "
                               "2 flop per DFMA); MEASURED_PEAKS.json has no fp64 entry",
                "kernel_ms": kernel_ms}
    if slots:
        achieved = 2.0 * slots * head.n / (kernel_ms * 1e-3) / 1e12
        roofline.update({"achieved": achieved, "frac": achieved / peak_tf,
                         "fp64_pipe_instructions_per_dof_step": slots, "instruction_count_source": slots_src,
                         "convention": "every FP64-pipe instruction (DFMA/DMUL/DADD) occupies one DFMA issue "
                                       "slot = 2 flop of the peak; frac is issue-slot utilisation"})
    else:
        roofline.update({"achieved": None, "frac": None})
    hbm_peak, hbm_src = None, None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    hbm_achieved = ALGO_BYTES[head.name] * head.n / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_slots.json")) as f:
            e = json.load(f).get(head.name, {})
        if e.get("ncu_dram_bytes_per_dof_step"):
            traffic = e["ncu_dram_bytes_per_dof_step"] * head.n
    except (OSError, ValueError):
        pass
    roofline["traffic"] = traffic
    roofline["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                       "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src,
                       "algorithmic_bytes_per_dof_step": ALGO_BYTES[head.name]}

    info = head.model.launch_info(args.block)

    # ------------------------------------------------ parity sample at full size (checker only)
    parity = None
    if args.parity_rows > 0 and args.scheme == "rk4":
        per = [dict(p.parity_sample(args.parity_rows), membrane_model=p.name) for p in parts]
        worst = max(max(q["max_rel_err_states"], q["max_rel_err_currents"]) for q in per)
        worst = dist.max(worst)
        parity = {"parts": per, "max_rel_err_all_ranks": worst, "ok": bool(worst < 1e-10)}
    for p in parts:
        p.model.close()

    # ------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu_baseline:
        v, threads, sample, _ = cpu_port_rate(args.workload, 4, 1, args.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "scheme": f"O1: RK4 x {N_SUB}, the scheme the GPU runs (apples to apples)",
               "reference_semantics_lsoda": run_cpu_lsoda(head.name, 1500, 20240611)}

    if dist.rank == 0:
        rhs_evals = 4 * N_SUB + 1 if args.scheme == "rk4" else dp45_steps["rhs_evals_per_dof_step"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": dist.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # `config` has exactly the keys of the reference arm's line (same workload, same scheme);
            # everything descriptive about this arm is under `details`
            "config": {"workload": args.workload,
                       "membrane_model": head.name if not strong else [p.name for p in parts],
                       "baseline_config": cfg_text,
                       "dofs_per_gpu": part_specs[0][1] if not strong else None,
                       "scheme": args.scheme, "n_sub": N_SUB if args.scheme == "rk4" else None,
                       "dt": head.dt, "l2": l2_note(part_specs)},
            "details": {"dofs_this_run_per_gpu": [p.n for p in parts], "dofs_total": int(total_dofs),
                        "dp45": dp45_steps, "rhs_evals_per_dof_step": rhs_evals,
                        "stimulus": "masked, x[0] < 20e-6 (~32 % of DOFs)",
                        "parallelism": f"{dist.world} x contiguous DOF ranges, no collective",
                        "rank0_cpu_affinity": f"{len(cpus)} CPUs local to GPU {dev}" if cpus else "unchanged",
                        "unread_inputs": args.unread_inputs,
                        "block": args.block or 128, "registers_per_thread": info["registers_per_thread"],
                        "blocks_per_sm": info["blocks_per_sm"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                    "d2h_bytes_per_step": int(d2h_all), "steps": e2e_steps,
                    "h2d_bytes_offered_per_step": int(dist.world * offered) if not strong else None,
                    "host_filled_bytes_per_step": int(filled),
                    "inputs_not_sent": kept_names, "outputs_filled_on_host": literal_names,
                    "ms_per_step_wall": e2e_ms_max / e2e_steps, "ms_per_step_device": dev_ms / e2e_steps,
                    "ms_per_step_wall_per_rank": e2e_per_rank, "ms_per_step_device_per_rank": e2e_dev_per_rank,
                    "last_step_ms": last, "link": link_frac,
                    "api": "MembraneModel.step_exchange (kem_step_io): the 7 input columns of one PDE step "
                           "from pinned host memory, fused step, 4 output columns back, chunk-pipelined; inputs "
                           "the right-hand side never reads follow config.unread_inputs, outputs the generated "
                           "code assigns a literal are filled on the host",
                    "dropin": dropin or None},
            "link_ceiling": link,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "sustained": sustained,
            "roofline": roofline,
            "parity_sample": parity,
            "rhs_evals_per_s": value * rhs_evals,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        _emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="hh_ideal_1e7")
    ap.add_argument("--dofs", type=float, default=0, help="override DOFs per GPU")
    ap.add_argument("--block", type=int, default=0, choices=[0, 64, 128, 256])
    ap.add_argument("--scheme", choices=["rk4", "dp45"], default="rk4",
                    help="rk4 = scheme O1 (the benchmark's normative scheme); dp45 = error-controlled O3")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample size")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-link-probe", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--unread-inputs", choices=["auto", "shadow", "upload", "discard"], default="discard",
                    help="policy for pushed PDE columns the right-hand side never reads (Cl_e, Cl_i)")
    ap.add_argument("--sustain-seconds", type=float, default=1.0,
                    help="after the K timed steps, repeat the resident loop for at least this long")
    ap.add_argument("--parity-rows", type=int, default=100_000,
                    help="random DOFs of the workload checked against the CPU oracle (0 = off)")
    args = ap.parse_args()
    # exactly one line on stdout: library chatter (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _emit
    _emit = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())     # noqa: E731
    # the reference arm is CPU work on rank 0 only: no process group, the other ranks just exit
    dist = Dist(init=args.impl != "reference")
    if dist.world != args.gpus and dist.world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={dist.world}")
    if args.gpus > 1 and dist.world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nnodes=1 "
                         "--nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...")
    try:
        if args.impl == "reference":
            run_reference(args, dist)
        else:
            run_gpu(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
