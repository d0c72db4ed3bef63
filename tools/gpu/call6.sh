# N=2: the exchange kernel's footprint next to the join (blocks per launch): 296 = 2 per SM, 148 = 1 per SM, 0 = as many as fit
mkdir -p gpurun_out
for B in 296 148; do
RWGPU_EXCHANGE_BLOCKS=$B BENCH_NO_VERIFY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((B % 10)) bench.py --gpus 2 --steps 20 --warmup 3 --legs value > gpurun_out/r2j_n2_b$B.json 2> gpurun_out/r2j_n2_b$B.err
python -c "import json; d=json.load(open('gpurun_out/r2j_n2_b$B.json')); print('BLOCKS $B', d['value']/1e9, d['ms_per_step'], d['roofline']['kernel_ms_avg'], d['host_ms_per_step'])"
done
