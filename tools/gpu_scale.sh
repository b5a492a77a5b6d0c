#!/usr/bin/env bash
# scaling check: bench.py under torchrun on N GPUs of one box (NCCL all-gather of the results)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
tail -1 gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n_gpus', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['clocks'])"
