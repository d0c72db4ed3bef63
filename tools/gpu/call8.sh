mkdir -p gpurun_out
timeout 330 python bench.py --steps 20 --warmup 3 --legs value,retract,hot,e2e,agg > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -2 gpurun_out/r2m_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2m_bench.json'))
print('value',d['value'],d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'verified',d.get('verified'))
for k in ('retract','hot','e2e','secondary','secondary_hot_keys','secondary_retract'):
    v=d.get(k)
    if isinstance(v,dict): print(k,{x:v[x] for x in v if x in ('value','ms_per_step','verified','ms_per_epoch','one_call_at_a_time')})
PY
