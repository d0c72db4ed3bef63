# round-2 GPU call 2f: all GPU tests, full bench, ncu launch list + full captures of the dominant kernels
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 1200 python -m pytest $f -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/r2f_$b.txt 2>&1
  echo "== $b: $(tail -1 gpurun_out/r2f_$b.txt)"
  grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2f_$b.txt | head -12
done
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench_full.json 2> gpurun_out/r2f_bench_full.err; tail -3 gpurun_out/r2f_bench_full.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_full.json'))
print('value',d['value'],d['ms_per_step'],'kernel',d['roofline']['kernel_ms_avg'],'frac',d['roofline']['frac'],'verified',d.get('verified'),'build',d.get('build_rows_per_s'),d.get('build_first_push_ms'))
for k in ('retract','hot','e2e','secondary','secondary_hot_keys','secondary_retract','q1','chain','generic_join','cpu_baseline'):
    v=d.get(k)
    if isinstance(v,dict): print(k,{x:v[x] for x in v if x in ('value','ms_per_step','verified','ms_per_epoch','one_call_at_a_time','degree_flip_step','launches_per_step')})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --legs value,retract,agg > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uni_hot_kernel -s 13 -c 2 -o gpurun_out/r2f_prof_hot python bench.py --steps 2 --warmup 3 --legs value > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_flush_kernel\|agg_apply_fast -s 8 -c 4 -o gpurun_out/r2f_prof_agg python bench.py --steps 2 --warmup 3 --legs agg > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uni_tail_kernel -s 16 -c 2 -o gpurun_out/r2f_prof_tail python bench.py --steps 2 --warmup 3 --legs value,retract > /dev/null 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err; tail -3 gpurun_out/r2f_ref.err; cut -c1-300 gpurun_out/r2f_ref.json
ls -la gpurun_out | tail -8
