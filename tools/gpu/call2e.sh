# round-2 GPU call 2e: pipelined retract diagnosis, A/B of prefetch / deferred link, tests, full bench
mkdir -p gpurun_out
DBG_PIPE=1 DBG_BUILD=10000000 DBG_BATCH=1048576 DBG_NBATCH=25 DBG_PAIRS=524288 DBG_STEPS=4 timeout 900 python tools/gpu/debug_retract.py > gpurun_out/r2e_dbg_pipe.txt 2>&1; tail -40 gpurun_out/r2e_dbg_pipe.txt
for fl in 0 1 2 3; do
  RWGPU_UNI_FLAGS=$fl BENCH_NO_VERIFY=1 timeout 300 python bench.py --steps 20 --warmup 3 --legs value > gpurun_out/r2e_bench_flags$fl.json 2> /dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2e_bench_flags$fl.json')); print('FLAGS $fl step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_avg'], 'launches', d['gpu_launches'])"
done
for f in tests/test_gpu_join.py tests/test_gpu_shuffle.py tests/test_gpu_agg.py; do
  b=$(basename $f .py)
  timeout 1200 python -m pytest $f -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2e_$b.txt 2>&1
  echo "== $b: $(tail -1 gpurun_out/r2e_$b.txt)"
  grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2e_$b.txt | head -12
done
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err; tail -3 gpurun_out/r2e_bench_full.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_full.json'))
print('value',d['value'],d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'verified',d.get('verified'),'build',d.get('build_rows_per_s'),d.get('build_first_push_ms'))
for k in ('retract','hot','e2e','secondary','secondary_hot_keys','secondary_retract','q1','chain','generic_join','cpu_baseline'):
    v=d.get(k)
    if isinstance(v,dict): print(k,{x:v[x] for x in v if x in ('value','ms_per_step','verified','ms_per_epoch','one_call_at_a_time','degree_flip_step','launches_per_step')})
PY
RWGPU_TRACE=1 timeout 300 python bench.py --steps 4 --warmup 3 --legs value,retract > /dev/null 2> gpurun_out/r2e_trace.err; grep -E "collect\]|uni_finish S=1" gpurun_out/r2e_trace.err | tail -12
