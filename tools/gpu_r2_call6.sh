#!/usr/bin/env bash
# round 2, call 6: f4 tests, chain kernel with 24 KB slots
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu.log | tail -12 | cut -c1-300
t() { env "$@" timeout 200 python tools/time_step.py ${CFG:-B} 2>&1 | tail -1; }
t A=0; t METRO_NO_CHAIN=1; t A=0
CFG=D t A=0; CFG=D t METRO_NO_CHAIN=1
CFG=E t A=0; CFG=E t METRO_NO_CHAIN=1
bash tools/gpu_prof.sh > /dev/null 2>&1; grep -E "roles|\+" gpurun_out/roles.txt | cut -c1-220
grep " us$" gpurun_out/roles.log | grep "+" 
