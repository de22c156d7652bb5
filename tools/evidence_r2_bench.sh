#!/bin/bash
# Round-2 bench lines of the FINAL build on one B200 (run through gpurun): the contract line, the
# reference arm, every other workload, scheme O3.  tools/make_bench_report.py r2 collects them
# (plus the multi-GPU lines gpurun_out/r2f_bench_n*.json) into profiles/r2_bench.md.
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 500 python bench.py --steps 20 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
for w in hh_test_1e6 calibration_1e7 hh_tissue_1e7 glial_tissue_1e7; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-dropin > gpurun_out/r2f_bench_$w.json 2> gpurun_out/r2f_bench_$w.err
done
for w in hh_ideal_1e7 hh_tissue_1e7 calibration_1e7 glial_tissue_1e7; do
  timeout 300 python bench.py --workload $w --scheme dp45 --steps 20 --no-cpu-baseline --no-dropin > gpurun_out/r2f_bench_${w}_dp45.json 2> gpurun_out/r2f_bench_${w}_dp45.err
done
timeout 600 python bench.py --workload tissue_1e8 --steps 5 --no-cpu-baseline > gpurun_out/r2f_bench_tissue_1e8_n1.json 2> gpurun_out/r2f_bench_tissue_1e8_n1.err
timeout 60 python tools/small_n_latency.py > gpurun_out/r2f_small_n.txt 2>&1
ls -la gpurun_out | grep r2f_
