"""Host-link ceilings of every GPU of a box, one rank per GPU (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/link_probe_ranks.py

Each rank first copies ALONE (the others wait at a barrier), then all ranks copy at the same time:
pinned 5 MB copies streaming through 480 MB of host memory per direction, both directions busy.
Rank 0 prints one JSON object: per-GPU rates alone and concurrent, their sums, and the CPU set
each rank was bound to -- what explains a per-rank asymmetry of the end-to-end exchange."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from knpemi_b200 import _cabi  # noqa: E402
from knpemi_b200.affinity import bind_to_device  # noqa: E402

rank, world, dev = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
cpus = [] if os.environ.get("KNPEMI_NO_AFFINITY") else bind_to_device(dev)
NB, REPS, SPAN = 5 << 20, 96, 480 << 20


def barrier():
    if world > 1:
        dist.barrier()


def gather(x):
    if world == 1:
        return [x]
    t = torch.zeros(world, dtype=torch.float64, device="cuda")
    t[rank] = x
    dist.all_reduce(t)
    return [round(float(v), 1) for v in t.tolist()]


alone = {"h2d": 0.0, "d2h": 0.0, "h2d_both": 0.0, "d2h_both": 0.0}
for r in range(world):
    barrier()
    if r == rank:
        alone["h2d"], _ = _cabi.link_probe(dev, NB, REPS, 0, SPAN)
        _, alone["d2h"] = _cabi.link_probe(dev, NB, 0, REPS, SPAN)
        alone["h2d_both"], alone["d2h_both"] = _cabi.link_probe(dev, NB, REPS, REPS, SPAN)
barrier()
conc_h, conc_d = _cabi.link_probe(dev, NB, 2 * REPS, 2 * REPS, SPAN)
barrier()
mix_h, mix_d = _cabi.link_probe(dev, NB, 80, 48, SPAN)        # the 5-in / 3-out mix of one HH exchange
out = {"copy_bytes": NB, "host_span_bytes": SPAN,
       "alone_h2d": gather(alone["h2d"]), "alone_d2h": gather(alone["d2h"]),
       "alone_both_h2d": gather(alone["h2d_both"]), "alone_both_d2h": gather(alone["d2h_both"]),
       "concurrent_both_h2d": gather(conc_h), "concurrent_both_d2h": gather(conc_d),
       "concurrent_mix_h2d": gather(mix_h), "concurrent_mix_d2h": gather(mix_d),
       "cpus_bound": gather(float(len(cpus)))}
if rank == 0:
    for k in ("alone_both", "concurrent_both", "concurrent_mix"):
        out[k + "_sum"] = round(sum(out[k + "_h2d"]) + sum(out[k + "_d2h"]), 1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
