mkdir -p gpurun_out
i=0
for cfg in "RWGPU_EXCHANGE_COOP=1" "BENCH_EX_PRIO=1" "RWGPU_EXCHANGE_COOP=1 RWGPU_EXCHANGE_BLOCKS=444"; do
i=$((i+1))
env $cfg BENCH_NO_VERIFY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$i bench.py --gpus 2 --steps 20 --warmup 3 --legs value > gpurun_out/r2k_n2_$i.json 2> gpurun_out/r2k_n2_$i.err
python -c "import json; d=json.load(open('gpurun_out/r2k_n2_$i.json')); print('$cfg', d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['host_ms_per_step'])" || tail -5 gpurun_out/r2k_n2_$i.err
done
