mkdir -p gpurun_out
timeout 300 python tools/gpu/bench_exchange.py > gpurun_out/r2h_exchange.txt 2>&1; tail -5 gpurun_out/r2h_exchange.txt
timeout 900 python -m pytest tests/test_gpu_shuffle.py tests/test_gpu_join.py tests/test_gpu_persistence.py -q -m gpu --timeout 500 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --legs value,retract,generic > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -2 gpurun_out/r2h_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print('value',d['value'],d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'verified',d.get('verified'))
for k in ('retract','generic_join'):
    v=d.get(k)
    if isinstance(v,dict): print(k,{x:v[x] for x in v if x in ('value','ms_per_step','verified','degree_flip_step','launches_per_step')})
PY
