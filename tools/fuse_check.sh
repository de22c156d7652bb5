#!/bin/bash
# GPU check of the shared-exponential build: parity suite, bench line, one full ncu capture.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -15 gpurun_out/gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.json
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 $B > gpurun_out/plain_b.log 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:kem_step_kernel -s 3 -c 1 \
      -f -o gpurun_out/prof_hh_ideal $B > gpurun_out/ncu_hh_ideal.log 2>&1
ls -la gpurun_out/prof_hh_ideal.ncu-rep
