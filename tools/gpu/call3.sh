# round-2 GPU call 3 (2 GPUs): the 2-rank exchange test, then the bench at N=2 (P2P exchange) and with the NCCL exchange
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_shuffle.py -q -m gpu -k two_gpus --timeout 500 -p no:cacheprovider > gpurun_out/r2c_two_gpu_test.txt 2>&1; tail -5 gpurun_out/r2c_two_gpu_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --legs value,e2e > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; tail -5 gpurun_out/r2c_bench_n2.err | cut -c1-300; cut -c1-900 gpurun_out/r2c_bench_n2.json
RWGPU_EXCHANGE=nccl NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 10 --warmup 3 --legs value > gpurun_out/r2c_bench_n2_nccl.json 2> gpurun_out/r2c_bench_n2_nccl.err; grep -m3 "comm\|NVLS\|nranks" gpurun_out/r2c_bench_n2_nccl.err | cut -c1-200; cut -c1-400 gpurun_out/r2c_bench_n2_nccl.json
