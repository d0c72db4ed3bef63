# round-2 GPU call 3 (N GPUs, default 2): the 2-rank exchange test, then the bench at N with the one-kernel exchange (default),
# the region exchange (p2p) and NCCL; N=1 on the same box for the scaling ratio
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_shuffle.py -q -m gpu -k two_gpus --timeout 500 -p no:cacheprovider > gpurun_out/r2_two_gpu_test.txt 2>&1; tail -5 gpurun_out/r2_two_gpu_test.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --legs value > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; cut -c1-300 gpurun_out/r2_scale_n1.json
for ex in flat p2p; do
  RWGPU_EXCHANGE=$ex timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --legs value,e2e > gpurun_out/r2_scale_n${N}_$ex.json 2> gpurun_out/r2_scale_n${N}_$ex.err; tail -4 gpurun_out/r2_scale_n${N}_$ex.err | cut -c1-300; cut -c1-700 gpurun_out/r2_scale_n${N}_$ex.json; echo
done
RWGPU_EXCHANGE=nccl NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --legs value > gpurun_out/r2_scale_n${N}_nccl.json 2> gpurun_out/r2_scale_n${N}_nccl.err; grep -m3 "comm\|NVLS\|nranks" gpurun_out/r2_scale_n${N}_nccl.err | cut -c1-200; cut -c1-400 gpurun_out/r2_scale_n${N}_nccl.json
