# round-2 GPU call: microbenchmark of the unified-table access pattern, GPU test suite, bench legs, CPU arm sweep
mkdir -p gpurun_out
(nproc; cat /sys/fs/cgroup/cpu.max 2>&1; lscpu | head -25; numactl -H 2>&1 | head -12; nvidia-smi topo -m 2>&1 | head -20) > gpurun_out/r2_host.txt 2>&1
timeout 300 ./build/ubench_unified > gpurun_out/r2_ubench_unified.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2_pytest_gpu.txt 2>&1; tail -40 gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --legs value,agg > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -5 gpurun_out/r2_bench_a.err; cut -c1-2500 gpurun_out/r2_bench_a.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_ref_a.json 2> gpurun_out/r2_ref_a.err; tail -6 gpurun_out/r2_ref_a.err
cat gpurun_out/r2_ubench_unified.txt
