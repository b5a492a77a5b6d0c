#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
METRO_ROLE_PROF=1 timeout 600 python bench.py --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
cat gpurun_out/bench_B.json
