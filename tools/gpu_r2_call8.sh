#!/usr/bin/env bash
# round 2, call 8: root kernel v2 (image pack folded in, paired conv rows), device-stamp profiling
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py tests/test_golden.py tests/test_crops.py -m gpu -q -x > gpurun_out/pytest_root2.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_root2.log | tail -15 | cut -c1-300
t() { env "$@" timeout 200 python tools/time_step.py ${CFG:-B} 2>&1 | tail -1; }
t A=0; t METRO_ROOT_V1=1; t A=0; t METRO_ROOT_V1=1
CFG=A t A=0; CFG=A t METRO_ROOT_V1=1
timeout 900 python bench.py --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_B.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']),
          'conv frac', round(d['roofline']['frac'], 3), 'extra', round(d['roofline']['extra_frac_step_minus_others'], 3), round(d['roofline']['extra_frac_conv_flops_over_whole_step'], 3),
          'step frac', round(d['details']['tensor_frac_whole_step'], 3), 'sam', round(d['roofline_softargmax']['achieved']),
          round(d['roofline_softargmax']['us_per_launch'], 2), 'us', d['clocks'], d['roofline']['other_ms'], d['roofline']['ms_sum_of_launches'])
except Exception as e:
    print('bench parse failed', e)
PY
grep " us$" gpurun_out/bench_B.err | awk '{printf "%-44s %7.1f   ", $1, $2; if (NR%3==0) printf "\n"} END{printf "\n"}'
METRO_PROFILE_EVENTS=1 timeout 300 python tools/ncu_targets.py roles 2>/dev/null | awk '{s+=$2} END{print "sum of event-timed launches", s, "us"}'
timeout 300 python tools/ncu_targets.py roles 2>/dev/null | awk '{s+=$2} END{print "sum of stamp-timed launches", s, "us"}'
