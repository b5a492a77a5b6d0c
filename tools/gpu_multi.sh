#!/usr/bin/env bash
# 2-GPU bench under torchrun (NCCL all-gather of the results) + the other BASELINE configs on one GPU
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"
tail -1 gpurun_out/bench_2gpu.json | cut -c1-700
for c in D C E A; do
  timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "config $c rc=$?"
  tail -1 gpurun_out/bench_$c.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], '| value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'conv frac', round(d['roofline']['frac'],3), 'sam GB/s', round(d['roofline_softargmax']['achieved']))" || tail -3 gpurun_out/bench_$c.err
done
