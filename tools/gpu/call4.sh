# round-2 scaling run on N GPUs of one box (NGPU=4 or 8): N=1 on the same box, then N with the one-kernel exchange
N=${NGPU:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --legs value > gpurun_out/r2_scale${N}_n1.json 2> gpurun_out/r2_scale${N}_n1.err; cut -c1-200 gpurun_out/r2_scale${N}_n1.json
for M in 2 $N; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 2953$M bench.py --gpus $M --steps 20 --warmup 3 --legs value,e2e > gpurun_out/r2_scale${N}_n$M.json 2> gpurun_out/r2_scale${N}_n$M.err; tail -3 gpurun_out/r2_scale${N}_n$M.err | cut -c1-300; cut -c1-260 gpurun_out/r2_scale${N}_n$M.json; echo
done
