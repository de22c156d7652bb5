"""The N > 1 host path on CPU: torchrun, world_size 2, gloo backend."""
import json
import os
import subprocess
import sys

import pytest

from knpemi_b200.sharding import dof_ranges

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script_args, port):
    env = dict(os.environ, KNPEMI_BENCH_BACKEND="gloo", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("n,parts", [(10, 1), (10, 3), (7, 8), (0, 2), (10_000_000, 8), (100_000_001, 8)])
def test_dof_ranges_partition(n, parts):
    r = dof_ranges(n, parts)
    assert len(r) == parts and r[0][0] == 0 and r[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all(e >= b for b, e in r)
    per = (n + parts - 1) // parts
    assert all(e - b == per for b, e in r[:-1] if e < n)


def test_sharded_ranks_reproduce_the_unsharded_result():
    r = _torchrun([os.path.join("tests", "native", "gloo_worker.py")], 29611)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GLOO_WORKER_OK" in r.stdout


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], 29612)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 2 and rec["unit"] == "DOF-steps/s"
    assert rec["cpu_baseline"]["kind"] == "port" and rec["e2e"]["h2d_bytes_per_step"] == 0
    assert rec["value"] > 0
