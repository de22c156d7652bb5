set -u
mkdir -p gpurun_out
for w in hh_tissue_1e7 calibration_1e7 hh_test_1e6; do
  timeout 100 python bench.py --workload $w --steps 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 100 python bench.py --scheme dp45 --steps 10 --no-cpu-baseline > gpurun_out/bench_hh_ideal_1e7_dp45.json 2> gpurun_out/bench_dp45.err
python - <<'PY'
import json,glob
for f in ("bench_hh_tissue_1e7","bench_calibration_1e7","bench_hh_test_1e6","bench_hh_ideal_1e7_dp45"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, "%.4e"%d["value"], d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d["roofline"]["frac"])
    except Exception as ex: print(f,"ERR",ex)
PY
