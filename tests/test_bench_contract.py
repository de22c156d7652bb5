"""The JSON line of `bench.py --impl reference` (the arm that needs no GPU): every key the
driver's contract names, the same `config` keys as the GPU arm writes, one line on stdout."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    env = dict(os.environ, KNPEMI_BENCH_REF_BUDGET_S="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", *extra], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys():
    d = _run()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "membrane DOF-steps/sec (fp64)" and d["unit"] == "DOF-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["scaling"] == "weak"
    assert set(d["config"]) == {"workload", "membrane_model", "baseline_config", "dofs_per_gpu", "scheme", "n_sub", "dt", "l2"}
    assert d["config"]["workload"] == "hh_ideal_1e7" and d["config"]["dofs_per_gpu"] == 10_000_000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "DOFs of hh_ideal" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_of_the_two_model_workload():
    d = _run("--workload", "tissue_1e8")
    assert d["scaling"] == "strong" and d["config"]["membrane_model"] == ["hh_tissue", "glial_tissue"]
    assert "hh_tissue" in d["cpu_baseline"]["sample"] and "glial_tissue" in d["cpu_baseline"]["sample"]


def test_gpu_arm_refuses_to_run_without_a_device():
    """No CPU fallback: the product arm fails loudly on a box without a GPU."""
    import pytest
    try:
        import ctypes
        ctypes.CDLL("libcuda.so.1")
        pytest.skip("a CUDA driver is present")
    except OSError:
        pass
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
