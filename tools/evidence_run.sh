#!/bin/bash
# Round evidence on one B200 (run through gpurun from the repo root): bench lines, the ncu
# launch list of the bench command and one `--set full` capture per model family.
# Every ncu command runs only after the identical plain command exited 0.
# About 12 GPU-minutes.  The multi-GPU lines are separate calls (charged N x):
#   gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#       --master-port 29611 bench.py --gpus N > gpurun_out/bench_nN.json'
# and tools/make_bench_report.py turns gpurun_out/bench_*.json into profiles/<round>_bench.md.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for w in hh_test_1e6 calibration_1e7 hh_tissue_1e7 glial_tissue_1e7; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
for w in hh_ideal_1e7 hh_tissue_1e7 calibration_1e7 glial_tissue_1e7; do
  timeout 300 python bench.py --workload $w --scheme dp45 --steps 20 --no-cpu-baseline > gpurun_out/bench_${w}_dp45.json 2> gpurun_out/bench_${w}_dp45.err
done
timeout 120 python tools/quick_dropin.py 1e7 --register > gpurun_out/dropin_register.txt 2>&1
timeout 300 $B > gpurun_out/plain_a.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 300 $B > gpurun_out/plain_b.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_kernel -s 3 -c 1 \
      -o gpurun_out/prof_hh_ideal $B > gpurun_out/ncu_hh_ideal.log 2>&1
for m in glial_tissue calibration hh_tissue; do
  Q="python tools/quick_perf.py $m 1e6 128"
  timeout 200 $Q > gpurun_out/plain_$m.log 2>&1 && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_kernel -s 3 -c 1 \
        -o gpurun_out/prof_$m $Q > gpurun_out/ncu_$m.log 2>&1
done
Q="python tools/quick_perf.py hh_ideal 1e6 128 dp45"
timeout 200 $Q > gpurun_out/plain_dp45.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_dp45 -s 3 -c 1 \
      -o gpurun_out/prof_hh_ideal_dp45 $Q > gpurun_out/ncu_dp45.log 2>&1
ls -la gpurun_out | tail -30
