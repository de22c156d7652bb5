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


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clock_sampler_has_samples_for_a_region_shorter_than_its_start_up(tmp_path, monkeypatch):
    """`nvidia-smi` needs a few hundred ms before its first row and samples every 50 ms; the default
    timed region is 0.15 s.  The `clocks` object must still carry samples (a line without them is
    rejected), from inside the region or, when it is shorter than a period, either side of it."""
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.4\nwhile true; do\n"
                    "echo '0, 1965, 1965, 300.5, Not Active, Not Active, Not Active, Active'; sleep 0.05; done\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    bench = _bench_module()
    for seconds in (0.15, 0.005):
        s = bench.ClockSampler(0)
        s.start()
        t0 = time.perf_counter()
        time.sleep(seconds)
        got = s.stop(t0, time.perf_counter())
        assert got["samples"] >= 1 and got["sm_mhz"] == 1965.0 and got["sm_max_mhz"] == 1965.0
        assert got["reasons"] == ["sw_power_cap"]


def test_clock_sampler_without_nvidia_smi_reports_no_samples(tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))                   # no nvidia-smi anywhere
    bench = _bench_module()
    s = bench.ClockSampler(0)
    s.start()
    assert s.stop(0.0, 1.0)["samples"] == 0
    dead = tmp_path / "nvidia-smi"                              # one that exits at once
    dead.write_text("#!/bin/sh\nexit 3\n")
    dead.chmod(0o755)
    s = bench.ClockSampler(0)
    s.start()
    assert s.stop(0.0, 1.0)["samples"] == 0
