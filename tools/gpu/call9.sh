mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider -x 2>&1 | tail -3
timeout 100 python bench.py --steps 20 --warmup 3 --legs q1,chain,generic,cpu > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -2 gpurun_out/r2n_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
for k in ('q1','chain','generic_join','cpu_baseline'):
    v=d.get(k)
    if isinstance(v,dict): print(k,{x:v[x] for x in v if x in ('value','ms_per_step','verified','degree_flip_step','launches_per_step')})
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
