"""Host-side staging copy pool of the runtime (csrc/kem_copy_pool.h), CPU only."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("threads", ["", "1", "4"])
def test_parallel_copies_are_exact_and_the_process_exits(tmp_path, threads):
    exe = tmp_path / "copy_pool_host"
    subprocess.run(["g++", "-O2", "-pthread", "-I", os.path.join(ROOT, "knp-emi-fenics-x_b200", "csrc"),
                    "-o", str(exe), os.path.join(ROOT, "tests", "native", "copy_pool_host.cpp")], check=True)
    env = dict(os.environ)
    if threads:
        env["KNPEMI_COPY_THREADS"] = threads
    r = subprocess.run([str(exe), "300"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and "COPY_POOL_OK" in r.stdout
