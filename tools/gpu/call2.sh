# round-2 GPU call 2: full GPU test suite (one pytest process per file: a sticky CUDA error cannot poison the others),
# bench (all legs), host timeline, occupancy variants of the hot join kernel, ncu launch list + full capture, CPU arm
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/r2b_$b.txt 2>&1
  echo "== $b: $(tail -1 gpurun_out/r2b_$b.txt)"
  grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2b_$b.txt | head -12
done
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err; tail -3 gpurun_out/r2b_bench_full.err; cut -c1-600 gpurun_out/r2b_bench_full.json
BENCH_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --legs value > gpurun_out/r2b_bench_trace.json 2> gpurun_out/r2b_bench_trace.err; grep -A12 "^\[trace\]" gpurun_out/r2b_bench_trace.err | head -14
for mb in 3 5 6; do
  RWGPU_UNI_MINB=$mb BENCH_NO_VERIFY=1 timeout 300 python bench.py --steps 10 --warmup 3 --legs value > gpurun_out/r2b_bench_minb$mb.json 2> /dev/null
  python -c "import json; d=json.load(open('gpurun_out/r2b_bench_minb$mb.json')); print('MINB $mb', d['ms_per_step'], d['roofline']['kernel_ms_avg'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 3 --legs value,agg > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:uni_hot_kernel -s 13 -c 2 -o gpurun_out/r2b_prof_hot python bench.py --steps 2 --warmup 3 --legs value > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:agg_flush_kernel\|agg_apply_fast -s 8 -c 4 -o gpurun_out/r2b_prof_agg python bench.py --steps 2 --warmup 3 --legs agg > /dev/null 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2b_ref.json 2> gpurun_out/r2b_ref.err; tail -6 gpurun_out/r2b_ref.err
ls -la gpurun_out | tail -20
