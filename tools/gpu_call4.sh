#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
cat gpurun_out/bench_B.json
for c in 32 128; do METRO_HOST_CHUNK=$c timeout 300 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c', d['e2e'], d['value'])"; done
