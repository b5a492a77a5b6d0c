#!/usr/bin/env bash
# round 2, call 5: chained conv3 + conv1 kernel (conv_chain.cu): correctness, then A/B timing
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "bit_identical or large_batch or sliced or graph" > gpurun_out/pytest_chain.log 2>&1; echo "pytest chain rc=$?"; grep -v "^$" gpurun_out/pytest_chain.log | tail -15 | cut -c1-300
t() { env "$@" timeout 200 python tools/time_step.py ${CFG:-B} 2>&1 | tail -1; }
t A=0; t METRO_NO_CHAIN=1; t A=0; t METRO_NO_DATAFLOW=1
CFG=D t A=0; CFG=D t METRO_NO_CHAIN=1
CFG=C t A=0; CFG=C t METRO_NO_CHAIN=1
CFG=E t A=0; CFG=E t METRO_NO_CHAIN=1
CFG=A t A=0; CFG=A t METRO_NO_CHAIN=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu.log | tail -8 | cut -c1-300
timeout 900 python bench.py --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_B.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']),
          'conv frac', round(d['roofline']['frac'], 3), 'extra', round(d['roofline']['extra_frac_step_minus_others'], 3),
          'step frac', round(d['details']['tensor_frac_whole_step'], 3), 'sam', round(d['roofline_softargmax']['achieved']),
          round(d['roofline_softargmax']['us_per_launch'], 2), 'us', d['clocks'])
except Exception as e:
    print('bench parse failed', e)
PY
grep " us$" gpurun_out/bench_B.err | awk '{printf "%-50s %8.1f\n", $1, $2}'
