#!/bin/bash
# One 8-GPU call: sweep the exchange variants with all ranks sharing the host
# (tools/exchange_scaling.py), pick the fastest chunk count / stream count, and run the bench
# lines with that choice exported (KNPEMI_IO_CHUNKS, KNPEMI_IO_H2D_STREAMS).
set -u
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29611 --nproc-per-node $N"
mkdir -p gpurun_out
timeout 400 $T tools/exchange_scaling.py 1e7 > gpurun_out/r2_xscale_n$N.txt 2> gpurun_out/r2_xscale_n$N.err
cat gpurun_out/r2_xscale_n$N.txt
eval $(python - <<PY
import json
best = {}
for line in open("gpurun_out/r2_xscale_n$N.txt"):
    try:
        d = json.loads(line)
    except ValueError:
        continue
    if "variant" in d:
        best[d["variant"]] = d["ms_max"]
def get(prefix):
    for k, v in best.items():
        if k.startswith(prefix):
            return v
    return 1e9
streams = 1 if get("1 H2D stream") < min(get("2 H2D streams"), get("default (")) else 2
chunks = min(((get("default ("), 16), (get("8 chunks"), 8), (get("4 chunks (20"), 4), (get("32 chunks"), 32)))[1]
print(f"export KNPEMI_IO_H2D_STREAMS={streams} KNPEMI_IO_CHUNKS={chunks}")
PY
)
echo "chosen: KNPEMI_IO_H2D_STREAMS=$KNPEMI_IO_H2D_STREAMS KNPEMI_IO_CHUNKS=$KNPEMI_IO_CHUNKS" | tee gpurun_out/r2_n${N}_choice.txt
timeout 400 $T bench.py --gpus $N --steps 20 --no-dropin > gpurun_out/r2_bench_n${N}_tuned.json 2> gpurun_out/r2_bench_n${N}_tuned.err
timeout 400 $T bench.py --gpus $N --workload tissue_1e8 --steps 20 --no-dropin > gpurun_out/r2_bench_tissue_n${N}_tuned.json 2> gpurun_out/r2_bench_tissue_n${N}_tuned.err
tail -2 gpurun_out/r2_bench_n${N}_tuned.err
