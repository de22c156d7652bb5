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
HBM and CUDA events on the launching stream; `e2e` goes through the public
MembraneModel API with pinned HOST buffers, host<->device copies inside the
timed region; `roofline` states the FP64-pipe issue-slot utilisation of the
fused kernel against a DFMA peak measured in the same run (MEASURED_PEAKS.json
has no fp64 entry) plus the HBM view; `cpu_baseline` is the oracle port of the
reference's stepping on the host cores (rank 0, N = 1 only).
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

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.06):
                continue
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
                "reasons": sorted(reasons), "samples": len(sm)}


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


def run_cpu_oracle(model_name: str, target_seconds: float, seed: int):
    """Oracle port of the reference stepping on the host cores: (DOF-steps/s, threads, sample)."""
    from oracle import cpu_oracle
    from workloads import SETUP, synthetic_tables
    threads = host_threads()
    cfg = SETUP[model_name]
    c_probe = 4000 * max(threads, 1)
    S, P, X, mask = synthetic_tables(model_name, c_probe, seed)
    P[mask, _stim_col(model_name)] = cfg["stim"]
    t0 = time.perf_counter()
    cpu_oracle.step(model_name, S, P, 0.0, cfg["dt"], N_SUB, threads)
    rate = c_probe / max(time.perf_counter() - t0, 1e-9)
    n_steps = 4
    n = int(min(max(rate * target_seconds / n_steps, c_probe), 4_000_000))
    S, P, X, mask = synthetic_tables(model_name, n, seed)
    P[mask, _stim_col(model_name)] = cfg["stim"]
    t, t0 = 0.0, time.perf_counter()
    for _ in range(n_steps):
        bad = cpu_oracle.step(model_name, S, P, t, cfg["dt"], N_SUB, threads)
        assert bad == 0
        t += cfg["dt"]
    el = time.perf_counter() - t0
    return n * n_steps / el, threads, f"{n} DOFs x {n_steps} PDE steps of {model_name} (RK4 x {N_SUB}), {el:.1f} s"


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
    from oracle import cpu_oracle
    from workloads import SETUP, synthetic_tables
    model_name, n_per_gpu, cfg_text = WORKLOADS[args.workload]
    cfg = SETUP[model_name]
    threads = host_threads()
    probe_n = 4000 * max(threads, 1)
    S, P, X, mask = synthetic_tables(model_name, probe_n, 20240611)
    t0 = time.perf_counter()
    cpu_oracle.step(model_name, S, P, 0.0, cfg["dt"], N_SUB, threads)
    rate = probe_n / max(time.perf_counter() - t0, 1e-9)
    budget = 150.0 / max(args.steps + args.warmup, 1)            # whole run within ~2.5 min
    n = int(min(max(rate * min(budget, 6.0), probe_n), 4_000_000))
    S, P, X, mask = synthetic_tables(model_name, n, 20240611)
    P[mask, _stim_col(model_name)] = cfg["stim"]
    t = 0.0
    for _ in range(args.warmup):
        cpu_oracle.step(model_name, S, P, t, cfg["dt"], N_SUB, threads)
        t += cfg["dt"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bad = cpu_oracle.step(model_name, S, P, t, cfg["dt"], N_SUB, threads)
        assert bad == 0
        t += cfg["dt"]
    el = time.perf_counter() - t0
    value = n * args.steps / el
    sample = f"{n} DOFs per step ({model_name}, RK4 x {N_SUB}) out of {n_per_gpu} per GPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "membrane_model": model_name, "baseline_config": cfg_text,
                   "dofs_per_gpu": n_per_gpu, "scheme": "rk4", "n_sub": N_SUB, "dt": cfg["dt"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU stepping; timed: oracle port of odeSolver.py:107-122 with the reference's "
                "RHS, OpenMP over rows, all host threads (numbalsoda/dolfinx not installable offline)",
    }
    _emit(json.dumps(line))


# ------------------------------------------------------------------------ GPU arm
def run_gpu(args, dist: Dist):
    from knpemi_b200 import _cabi
    from knpemi_b200.ducks import PointSpace
    from knpemi_b200.odeSolver import MembraneModel
    from workloads import SETUP, builtin, load_tables, synthetic_tables

    model_name, n, cfg_text = WORKLOADS[args.workload]
    if args.dofs:
        n = int(args.dofs)
    cfg = SETUP[model_name]
    ode = builtin(model_name)
    dev = dist.local_rank
    if _cabi.device_count() <= dev:
        raise SystemExit("bench.py: no CUDA device for this rank -- the product has no CPU path")
    # several ranks on a multi-socket host: keep this rank's pinned buffers next to its GPU
    cpus = []
    if dist.world > 1 and not os.environ.get("KNPEMI_NO_AFFINITY"):
        from knpemi_b200.affinity import bind_to_device
        cpus = bind_to_device(dev)

    S, P, X, mask = synthetic_tables(model_name, n, seed=20240611 + dist.rank)
    model = MembraneModel(ode, None, 1, PointSpace(X), devices=[dev], verbose=False, n_sub=N_SUB,
                          block=args.block, scheme=args.scheme)
    load_tables(model, S, P)
    stim = {"stim_amplitude": cfg["stim"]}
    locator = lambda x: x[0] < 20e-6                 # noqa: E731  (run_2D.py:264)
    dt = cfg["dt"]

    # ------------------------------------------------ resident: inputs already in HBM
    for _ in range(max(args.warmup, 3)):
        model.step_async(dt, stim, locator)
    model.synchronize()
    sampler = ClockSampler(dev)
    sampler.start()
    time.sleep(0.12)
    dist.barrier()
    model.synchronize()
    launches0 = model.launch_count()
    wall0 = time.perf_counter()
    model.timer_begin()
    for _ in range(args.steps):
        model.step_async(dt, stim, locator)
    ms = model.timer_end()
    model.synchronize()
    wall1 = time.perf_counter()
    dist.barrier()
    launches = model.launch_count() - launches0
    clocks = sampler.stop(wall0, wall1)
    dp45_steps = None
    if args.scheme == "dp45":
        model.step_stats()                                   # discard warm-up counts
        for _ in range(3):
            model.step_async(dt, stim, locator)
        acc, rej = model.step_stats()
        dp45_steps = {"accepted_per_dof_step": acc / (3.0 * n), "rejected_per_dof_step": rej / (3.0 * n),
                      "rhs_evals_per_dof_step": 6.0 * (acc + rej) / (3.0 * n) + 1.0,
                      "rtol": model.rtol, "atol": model.atol}
    ms_max = dist.max(ms)
    total_dofs = dist.sum(float(n))
    value = total_dofs * args.steps / (ms_max * 1e-3)
    ms_per_step = ms_max / args.steps

    # ------------------------------------------------ end to end: host buffers through the API
    in_names, v_io, out_names = IO_COLUMNS[model_name]
    keep = []

    def pinned(src=None):
        pa = _cabi.PinnedArray(n)
        keep.append(pa)
        pa.array[:] = 0.0 if src is None else src
        return pa.array

    ins = {("parameter", k): pinned(P[:, ode.parameter_indices(k)]) for k in in_names}
    outs = {("parameter", k): pinned() for k in out_names}
    if v_io:
        v_in, v_out = pinned(np.asarray(model.states[:, ode.state_indices("V")])), pinned()
        ins[("state", "V")] = v_in
        outs[("state", "V")] = v_out
    h2d_offered = 8 * n * len(ins)
    d2h = 8 * n * len(outs)
    e2e_steps = max(min(args.steps, 20), 3)
    for _ in range(2):
        model.step_exchange(dt, ins, outs, stim, locator)
    dist.barrier()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tm = model.step_exchange(dt, ins, outs, stim, locator)
        dev_ms += tm["ms_total"]
        if v_io:                                      # the PDE side would hand phi_M back
            ins[("state", "V")], outs[("state", "V")] = outs[("state", "V")], ins[("state", "V")]
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    # bytes that really crossed the link: inputs to slots the right-hand side never reads
    # (Cl_e, Cl_i for the HH models) are kept in a host shadow by the library
    host_only = [k for (what, k) in ins if what == "parameter" and model.column_location(what, k) == "host"]
    h2d = 8 * n * (len(ins) - len(host_only))
    dist.barrier()
    e2e_ms_max = dist.max(e2e_wall_ms)
    e2e_per_rank = [round(v / e2e_steps, 3) for v in dist.gather(e2e_wall_ms)]
    e2e_dev_per_rank = [round(v / e2e_steps, 3) for v in dist.gather(dev_ms)]
    e2e_value = total_dofs * e2e_steps / (e2e_ms_max * 1e-3)
    last = dict(model.last_step_times)

    # ------------------------------------------------ the reference's own call sequence, unmodified:
    # 7 setter calls + step_lsoda + 4 getter calls per PDE step on pageable NumPy arrays
    # (utils.py:227-233, run_2D.py:98-109) -- what a user sees without touching solve_odes
    from knpemi_b200.ducks import ArrayFunction
    calls = None
    if v_io and in_names:
        u_in = {k: ArrayFunction(P[:, ode.parameter_indices(k)].copy()) for k in in_names}
        u_phi = ArrayFunction(np.asarray(model.states[:, ode.state_indices("V")]))
        u_out = {k: ArrayFunction(n) for k in out_names}

        def one_pde_step():
            for k, u in u_in.items():
                model.set_parameter(k, u)
            model.set_membrane_potential(u_phi)
            model.step_lsoda(dt, stim, locator)
            model.get_membrane_potential(u_phi)
            for k, u in u_out.items():
                model.get_parameter(k, u)

        one_pde_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            one_pde_step()
        call_ms = dist.max((time.perf_counter() - t0) * 1e3) / 3
        calls = {"value": total_dofs / (call_ms * 1e-3), "unit": UNIT, "ms_per_step_wall": call_ms,
                 "api": "set_parameter x6 + set_membrane_potential + step_lsoda + get_membrane_potential "
                        "+ get_parameter x3, pageable host arrays (staged through pinned buffers)"}

    # ------------------------------------------------ roofline of the fused kernel
    peak_tf, _ = _cabi.fp64_peak(dev)
    slots, slots_src = fp64_slots(model_name, N_SUB)
    if dp45_steps and slots:
        # O3 spends a data-dependent number of RHS evaluations: scale the per-RHS count of the
        # O1 kernel (same generated right-hand side) by the measured evaluations per DOF-step
        slots = slots / (4 * N_SUB + 1) * dp45_steps["rhs_evals_per_dof_step"]
        slots_src += " x measured RHS evaluations of scheme O3 (excludes controller arithmetic)"
    kernel_ms = ms / args.steps                       # this rank's average launch duration
    roofline = {"bound": "fp64", "unit": "TFLOP/s", "peak": peak_tf,
                "peak_source": "measured in this run: kem_fp64_peak (8 independent DFMA chains/thread, "
                               "2 flop per DFMA); MEASURED_PEAKS.json has no fp64 entry",
                "kernel_ms": kernel_ms}
    if slots:
        achieved = 2.0 * slots * n / (kernel_ms * 1e-3) / 1e12
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
    hbm_achieved = ALGO_BYTES[model_name] * n / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_slots.json")) as f:
            e = json.load(f).get(model_name, {})
        if e.get("ncu_dram_bytes_per_dof_step"):
            traffic = e["ncu_dram_bytes_per_dof_step"] * n
    except (OSError, ValueError):
        pass
    roofline["traffic"] = traffic
    roofline["hbm"] = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                       "frac": hbm_achieved / hbm_peak, "peak_source": hbm_src,
                       "algorithmic_bytes_per_dof_step": ALGO_BYTES[model_name]}

    info = model.launch_info(args.block)
    model.close()

    # ------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu_baseline:
        v, threads, sample = run_cpu_oracle(model_name, args.cpu_seconds, 20240611)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "scheme": f"O1: RK4 x {N_SUB}, the scheme the GPU runs (apples to apples)",
               "reference_semantics_lsoda": run_cpu_lsoda(model_name, 1500, 20240611)}

    if dist.rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": dist.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "membrane_model": model_name, "baseline_config": cfg_text,
                       "dofs_per_gpu": n, "scheme": args.scheme, "n_sub": N_SUB if args.scheme == "rk4" else None,
                       "dt": dt, "dp45": dp45_steps,
                       "rhs_evals_per_dof_step": 4 * N_SUB + 1 if args.scheme == "rk4"
                       else dp45_steps["rhs_evals_per_dof_step"], "stimulus": "masked, x[0] < 20e-6 (~32 % of DOFs)",
                       "l2": (f"inputs {ALGO_BYTES[model_name] * n / 1e6:.0f} MB per GPU > 126 MB L2, no flush needed"
                              if ALGO_BYTES[model_name] * n > 126e6 else
                              f"inputs {ALGO_BYTES[model_name] * n / 1e6:.0f} MB per GPU < 126 MB L2 and NOT "
                              "flushed: informational workload, not a contract line (compute-bound kernel)"),
                       "parallelism": f"{dist.world} x contiguous DOF ranges, no collective",
                       "rank0_cpu_affinity": f"{len(cpus)} CPUs local to GPU {dev}" if cpus else "unchanged",
                       "block": args.block or 128, "registers_per_thread": info["registers_per_thread"],
                       "blocks_per_sm": info["blocks_per_sm"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d * dist.world),
                    "d2h_bytes_per_step": int(d2h * dist.world), "steps": e2e_steps,
                    "h2d_bytes_offered_per_step": int(h2d_offered * dist.world),
                    "inputs_kept_in_host_shadow": host_only,
                    "ms_per_step_wall": e2e_ms_max / e2e_steps, "ms_per_step_device": dev_ms / e2e_steps,
                    "ms_per_step_wall_per_rank": e2e_per_rank, "ms_per_step_device_per_rank": e2e_dev_per_rank,
                    "last_step_ms": last,
                    "api": "MembraneModel.step_exchange (kem_step_io): the 7 input columns of one PDE step "
                           "from pinned host memory (those the right-hand side never reads stay in a host "
                           "shadow), fused step, 4 output columns back, chunk-pipelined",
                    "unmodified_reference_calls": calls},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "rhs_evals_per_s": value * (4 * N_SUB + 1 if args.scheme == "rk4"
                                        else dp45_steps["rhs_evals_per_dof_step"]),
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
