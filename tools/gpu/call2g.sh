mkdir -p gpurun_out
timeout 300 python tools/gpu/bench_exchange.py > gpurun_out/r2g_exchange.txt 2>&1; tail -8 gpurun_out/r2g_exchange.txt
timeout 600 python tools/gpu/debug_generic.py > gpurun_out/r2g_generic.txt 2>&1; tail -16 gpurun_out/r2g_generic.txt | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_shuffle.py tests/test_gpu_join.py -q -m gpu --timeout 500 -p no:cacheprovider -x 2>&1 | tail -3
