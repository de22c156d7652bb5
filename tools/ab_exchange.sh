#!/bin/bash
# one-GPU A/B of the exchange pipeline knobs (kem_step_io): copy threads, chunks, shadow
set -u
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 120 $B > gpurun_out/abx_$name.json 2> gpurun_out/abx_$name.err; }
run default X=1
run threads2 KNPEMI_COPY_THREADS=2
run threads4 KNPEMI_COPY_THREADS=4
run noshadow KNPEMI_NO_HOST_SHADOW=1
run chunks32 KNPEMI_IO_CHUNKS=32
run chunks8 KNPEMI_IO_CHUNKS=8
python - <<'PY'
import json
for k in ("default","threads2","threads4","noshadow","chunks32","chunks8"):
    try:
        d=json.loads(open(f"gpurun_out/abx_{k}.json").read().strip().splitlines()[-1]); e=d["e2e"]
        print(k, "value %.3e"%d["value"], "e2e %.3e"%e["value"], e["ms_per_step_wall"], e["last_step_ms"])
    except Exception as ex: print(k, "ERR", ex)
PY
