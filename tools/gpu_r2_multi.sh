#!/usr/bin/env bash
# multi-GPU evidence: `gpurun --gpus N -- bash tools/gpu_r2_multi.sh N [configs]`
N=${1:-2}; CFGS=${2:-B}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
if [ "$N" = 2 ]; then
  timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k two_devices > gpurun_out/pytest_2dev.log 2>&1; echo "two-device test rc=$?"; tail -2 gpurun_out/pytest_2dev.log
fi
for c in $CFGS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${c}_${N}gpu.json 2> gpurun_out/bench_${c}_${N}gpu.err; echo "bench $c x$N rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_${c}_${N}gpu.json') if l.startswith('{')][-1])
    print('${c} x${N}: value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), d['config'], d['details'].get('host_numa_binding'))
except Exception as e:
    print('parse failed', e)
PY
done
