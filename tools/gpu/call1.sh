mkdir -p gpurun_out
(nproc; cat /sys/fs/cgroup/cpu.max 2>&1; lscpu | head -25; numactl -H 2>&1 | head -12; nvidia-smi topo -m 2>&1 | head -20) > gpurun_out/r2_host.txt 2>&1
./build/ubench_unified > gpurun_out/r2_ubench_unified.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_join.py -x -q -m gpu > gpurun_out/r2_pytest_join.txt 2>&1; tail -15 gpurun_out/r2_pytest_join.txt
timeout 300 python bench.py --steps 10 --warmup 3 --legs value > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -3 gpurun_out/r2_bench_a.err; cut -c1-1800 gpurun_out/r2_bench_a.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_ref_a.json 2> gpurun_out/r2_ref_a.err; tail -6 gpurun_out/r2_ref_a.err
cat gpurun_out/r2_ubench_unified.txt
