# scaling point: N GPUs, the default (one-kernel, staged) exchange, value leg only
N=${NGPU:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 3 --legs value > gpurun_out/r2i_scale_n$N.json 2> gpurun_out/r2i_scale_n$N.err; tail -2 gpurun_out/r2i_scale_n$N.err | cut -c1-300; cut -c1-260 gpurun_out/r2i_scale_n$N.json; echo
