"""world_size-2 CPU worker (launched by tests/test_multirank_gloo.py under torchrun).

Each rank advances its contiguous DOF range of one workload with the CPU oracle (the
checker stands in for the device so the host-side range / gather / timing-reduction logic
runs without a GPU); rank 0 checks the gathered result bitwise against the unsharded run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "knp-emi-fenics-x_b200"), ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

os.environ["KNPEMI_BENCH_BACKEND"] = "gloo"
import bench  # noqa: E402
from knpemi_b200.sharding import dof_ranges, rank_range  # noqa: E402
from oracle import cpu_oracle  # noqa: E402
from workloads import SETUP, synthetic_tables  # noqa: E402


def main():
    d = bench.Dist()
    assert d.world == 2 and d.backend == "gloo"
    name, n = "hh_tissue", 10007
    S, P, X, mask = synthetic_tables(name, n, seed=1)
    P[mask, 8] = SETUP[name]["stim"]
    b, e = rank_range(n, d.rank, d.world)
    assert dof_ranges(n, 2) == [(0, 5004), (5004, 10007)]
    Sl, Pl = S[b:e].copy(), P[b:e].copy()
    t = 0.0
    for _ in range(2):
        assert cpu_oracle.step(name, Sl, Pl, t, 0.1, 25, 1) == 0
        t += 0.1
    d.barrier()
    # timing reduction used by bench.py: max over ranks, sum of DOFs
    assert d.max(float(d.rank + 1)) == 2.0
    assert d.sum(float(e - b)) == float(n)
    parts = [None, None]
    dist.all_gather_object(parts, (b, e, Sl, Pl))
    if d.rank == 0:
        Sg = np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])])
        Pg = np.concatenate([p[3] for p in sorted(parts, key=lambda p: p[0])])
        t = 0.0
        for _ in range(2):
            cpu_oracle.step(name, S, P, t, 0.1, 25, 1)
            t += 0.1
        assert np.array_equal(Sg, S) and np.array_equal(Pg, P)
        print("GLOO_WORKER_OK", flush=True)
    d.close()


if __name__ == "__main__":
    main()
