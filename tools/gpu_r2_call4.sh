#!/usr/bin/env bash
# round 2, call 4: tile-order / xform-width / graph-executor / early-prefetch experiments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_net_gpu.py tests/test_softargmax_gpu.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu.log | tail -12 | cut -c1-400
t() { env "$@" timeout 200 python tools/time_step.py ${CFG:-B} 2>&1 | tail -1; }
t A=0; t METRO_DF_KEEP_ALTERNATE=1; t METRO_NO_DATAFLOW=1; t METRO_XFORM_MAX_CB=128; t METRO_XFORM_MAX_CB=128 METRO_NO_DATAFLOW=1; t A=0
CFG=D t A=0; CFG=D t METRO_DF_KEEP_ALTERNATE=1; CFG=D t METRO_NO_DATAFLOW=1; CFG=D t METRO_XFORM_MAX_CB=128
CFG=C t A=0; CFG=C t METRO_NO_DATAFLOW=1; CFG=C t METRO_XFORM_MAX_CB=128
CFG=E t A=0; CFG=E t METRO_NO_DATAFLOW=1; CFG=E t METRO_XFORM_MAX_CB=128
CFG=A t A=0; CFG=A t METRO_GRAPH_MAX_BATCH=0; CFG=A t METRO_NO_DATAFLOW=1
timeout 300 python tools/sam_sweep.py all "0,0,0" > gpurun_out/sam_sweep_early.log 2>&1; grep f32 gpurun_out/sam_sweep_early.log
METRO_SAM_NO_EARLY=1 timeout 300 python tools/sam_sweep.py quick "0,0,0" > gpurun_out/sam_sweep_noearly.log 2>&1; grep f32 gpurun_out/sam_sweep_noearly.log
