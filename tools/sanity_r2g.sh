python tools/small_n_latency.py > gpurun_out/r2g_small_n.txt 2>&1
timeout 225 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
tail -3 gpurun_out/r2g_tests.log; cat gpurun_out/r2g_small_n.txt
