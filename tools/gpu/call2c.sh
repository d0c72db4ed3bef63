# round-2 GPU call 2c: diagnose the retract leg (row-level check at two scales) and the slow build (host timeline)
mkdir -p gpurun_out
timeout 600 python tools/gpu/debug_retract.py > gpurun_out/r2c_dbg_small.txt 2>&1; tail -30 gpurun_out/r2c_dbg_small.txt
DBG_BUILD=4000000 DBG_BATCH=1048576 DBG_NBATCH=10 DBG_PAIRS=524288 DBG_STEPS=4 RWGPU_TRACE=1 timeout 900 python tools/gpu/debug_retract.py > gpurun_out/r2c_dbg_large.txt 2>&1; grep -v "^\s*\[push_dev" gpurun_out/r2c_dbg_large.txt | tail -60
timeout 900 python -m pytest tests/test_gpu_join.py -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2c_test_gpu_join.txt 2>&1
echo "== test_gpu_join: $(tail -1 gpurun_out/r2c_test_gpu_join.txt)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2c_test_gpu_join.txt | head -30
timeout 600 python -m pytest tests/test_gpu_persistence.py tests/test_gpu_filter.py -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2c_test_misc.txt 2>&1
echo "== persistence+filter: $(tail -1 gpurun_out/r2c_test_misc.txt)"; grep -E "^(FAILED|ERROR)|^E  " gpurun_out/r2c_test_misc.txt | head -30
RWGPU_TRACE=1 BENCH_NO_VERIFY=1 timeout 600 python bench.py --steps 4 --warmup 3 --legs value,retract > gpurun_out/r2c_bench_trace.json 2> gpurun_out/r2c_bench_trace.err
grep -E "uni_enqueue S=1|uni_finish S=1" gpurun_out/r2c_bench_trace.err | head -60
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_trace.json')); print(d['build_rows_per_s'], d['ms_per_step'], d.get('retract',{}).get('ms_per_step'))"
timeout 600 python bench.py --steps 10 --warmup 3 --legs e2e > gpurun_out/r2c_bench_e2e.json 2> gpurun_out/r2c_bench_e2e.err; tail -3 gpurun_out/r2c_bench_e2e.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_e2e.json')); print(d.get('e2e'))"
