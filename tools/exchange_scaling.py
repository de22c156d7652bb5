"""What the pipelined exchange loses when N ranks share one host (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/exchange_scaling.py

Every rank owns 1e7 hh_ideal DOFs on its GPU and runs the same 5-in / 3-out (+1 filled) exchange
in several variants, all ranks starting together; rank 0 prints one JSON line per variant with the
per-rank mean time of an exchange and the aggregate bytes/s over the host link.  The last lines
are the link probe with the same copies (no kernel, no dependencies), five single shots."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from knpemi_b200 import _cabi  # noqa: E402
from knpemi_b200.affinity import bind_to_device  # noqa: E402
from knpemi_b200.ducks import PointSpace  # noqa: E402
from knpemi_b200.odeSolver import MembraneModel  # noqa: E402
from workloads import SETUP, builtin, load_tables, synthetic_tables  # noqa: E402

rank, world, dev = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
if not os.environ.get("KNPEMI_NO_AFFINITY"):
    bind_to_device(dev)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
name = "hh_ideal"
ode = builtin(name)
S, P, X, mask = synthetic_tables(name, n, seed=20240611 + rank)
m = MembraneModel(ode, None, 1, PointSpace(X), devices=[dev], verbose=False, unread_inputs="discard")
load_tables(m, S, P)
stim, loc, dt = {"stim_amplitude": SETUP[name]["stim"]}, (lambda x: x[0] < 20e-6), SETUP[name]["dt"]


def pinned(src=None):
    a = _cabi.pinned_empty(n)
    a[:] = 0.0 if src is None else src
    return a


ins = {("parameter", k): pinned(P[:, ode.parameter_indices(k)]) for k in ("K_e", "K_i", "Na_e", "Na_i", "Cl_e", "Cl_i")}
ins[("state", "V")] = pinned(S[:, 3])
outs = {("state", "V"): pinned(), **{("parameter", k): pinned() for k in ("I_ch_Na", "I_ch_K", "I_ch_Cl")}}
outs_nofill = {k: v for k, v in outs.items() if k != ("parameter", "I_ch_Cl")}
BYTES = 8 * n * (5 + 3)


def barrier():
    if world > 1:
        dist.barrier()


def gather(x):
    if world == 1:
        return [round(x, 2)]
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[rank] = x
    dist.all_reduce(t)
    return [round(float(v), 2) for v in t.tolist()]


def variant(label, chunks=0, streams=0, o=outs, n_sub=None, lockstep=False, reps=10):
    _cabi.check(m._lib.kem_set_io_tuning(m._h, chunks, streams), "kem_set_io_tuning")
    for _ in range(2):
        m.step_exchange(dt, ins, o, stim, loc, n_sub=n_sub)
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(reps):
        if lockstep:
            barrier()
        dev_ms += m.step_exchange(dt, ins, o, stim, loc, n_sub=n_sub)["ms_total"]
    wall = (time.perf_counter() - t0) * 1e3 / reps
    per_rank, dev_rank = gather(wall), gather(dev_ms / reps)
    barrier()
    if rank == 0:
        agg = sum(BYTES / (ms * 1e-3) for ms in per_rank) / 1e9
        print(json.dumps({"variant": label, "ms_wall_per_rank": per_rank, "ms_device_per_rank": dev_rank,
                          "ms_max": max(per_rank), "aggregate_link_gbs": round(agg, 1),
                          "dof_steps_per_s": round(world * n / (max(per_rank) * 1e-3), 0)}), flush=True)


variant("default (2 H2D streams, 16 chunks)")
variant("1 H2D stream", streams=1)
variant("2 H2D streams", streams=2)
variant("8 chunks (10 MB copies)", chunks=8)
variant("4 chunks (20 MB copies)", chunks=4)
variant("4 chunks, 1 H2D stream", chunks=4, streams=1)
variant("32 chunks", chunks=32)
variant("no I_ch_Cl requested (no host fill)", o=outs_nofill)
variant("RK4 x 1 (no kernel work)", n_sub=1)
variant("lock-step ranks (barrier before every exchange)", lockstep=True)
variant("default again")
_cabi.check(m._lib.kem_set_io_tuning(m._h, 0, 0), "kem_set_io_tuning")
chunk_bytes = 8 * ((n + 15) // 16)
shots = []
for _ in range(5):
    barrier()
    h, d = _cabi.link_probe(dev, chunk_bytes, 80, 48, 480 << 20)
    ms = max(chunk_bytes * 80 / (h * 1e6), chunk_bytes * 48 / (d * 1e6))
    shots.append(gather(ms))
if rank == 0:
    print(json.dumps({"probe_same_copies_ms_per_rank_five_shots": shots,
                      "ms_max_per_shot": [max(s) for s in shots]}), flush=True)
m.close()
if world > 1:
    dist.destroy_process_group()
