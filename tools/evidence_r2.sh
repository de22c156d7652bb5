#!/bin/bash
# Round-2 ncu evidence on one B200 (run through gpurun from the repo root): the launch list of
# the bench command and one `--set full` capture per kernel family of the FINAL build.
# Every ncu command runs only after the identical plain command exited 0 (B200_PROFILING.md).
# About 8 GPU-minutes.  tools/ncu_summary.py turns the .ncu-rep files into profiles/r2_ncu_*.txt.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --no-link-probe --sustain-seconds 0 --parity-rows 0"
timeout 300 $B > gpurun_out/r2_plain_a.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
timeout 300 $B > gpurun_out/r2_plain_b.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_kernel -s 3 -c 1 \
      -o gpurun_out/r2_prof_hh_ideal $B > gpurun_out/r2_ncu_hh_ideal.log 2>&1
for m in hh_tissue glial_tissue calibration; do
  Q="python tools/quick_perf.py $m 1e6 128"
  timeout 200 $Q > gpurun_out/r2_plain_$m.log 2>&1 && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_kernel -s 3 -c 1 \
        -o gpurun_out/r2_prof_$m $Q > gpurun_out/r2_ncu_$m.log 2>&1
done
for m in hh_ideal calibration; do
  Q="python tools/quick_perf.py $m 1e6 128 dp45"
  timeout 200 $Q > gpurun_out/r2_plain_dp45_$m.log 2>&1 && \
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:kem_step_dp45 -s 3 -c 1 \
        -o gpurun_out/r2_prof_dp45_$m $Q > gpurun_out/r2_ncu_dp45_$m.log 2>&1
done
ls -la gpurun_out | grep r2_prof
