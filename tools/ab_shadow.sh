set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline"
timeout 240 $TR > gpurun_out/ab_shadow.json 2> gpurun_out/ab_shadow.err
KNPEMI_NO_HOST_SHADOW=1 timeout 240 $TR > gpurun_out/ab_noshadow.json 2> gpurun_out/ab_noshadow.err
KNPEMI_COPY_THREADS=2 timeout 240 $TR > gpurun_out/ab_shadow_t2.json 2> gpurun_out/ab_shadow_t2.err
timeout 240 $TR > gpurun_out/ab_shadow_again.json 2> gpurun_out/ab_shadow_again.err
python - <<'PY'
import json
for k in ("shadow","noshadow","shadow_t2","shadow_again"):
    try:
        d=json.loads(open(f"gpurun_out/ab_{k}.json").read().strip().splitlines()[-1])
        e=d["e2e"]; print(k, "value %.3e"%d["value"], "e2e %.3e"%e["value"], e.get("ms_per_step_ranks"), "unmod %.3e"%e["unmodified_reference_calls"]["value"])
    except Exception as ex: print(k, "ERR", ex)
PY
