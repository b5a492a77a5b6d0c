#!/usr/bin/env bash
# round 2, first GPU call: full parity suite, bench line, stem-chunk experiment, compute-sanitizer passes
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.json
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
grep -h "^parity" gpurun_out/pytest_gpu.log | head
timeout 900 python bench.py --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_B.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']),
          'conv frac', round(d['roofline']['frac'], 3), 'extra', round(d['roofline']['extra_frac_step_minus_others'], 3),
          'step frac', round(d['details']['tensor_frac_whole_step'], 3), 'sam', round(d['roofline_softargmax']['achieved']),
          round(d['roofline_softargmax']['us_per_launch'], 2), 'us', d['clocks'], d['details']['extended_region'])
except Exception as e:
    print('bench parse failed', e)
PY
for c in 0 16 32 48 64 128; do METRO_STEM_CHUNK=$c timeout 200 python tools/time_step.py B 2>&1 | tail -1; done
METRO_NO_ALTERNATE=1 timeout 200 python tools/time_step.py B 2>&1 | tail -1
for c in 0 16 32; do METRO_STEM_CHUNK=$c timeout 200 python tools/time_step.py D 2>&1 | tail -1; done
# compute-sanitizer (SURVEY section 5 "race detection"): small shapes, every kernel family
CS="compute-sanitizer --print-limit 20"
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool python tools/ncu_targets.py net 3 1 > gpurun_out/sanitizer_${tool}_net.log 2>&1; echo "$tool net rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_net.log
  timeout 600 $CS --tool $tool python -m pytest tests/test_softargmax_gpu.py -q -x -k "cross_cta or known or parity_vs" > gpurun_out/sanitizer_${tool}_sam.log 2>&1; echo "$tool sam rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_sam.log
  timeout 600 $CS --tool $tool python -m pytest tests/test_conv_gpu.py -q -x -k "residual or projection or float32" > gpurun_out/sanitizer_${tool}_conv.log 2>&1; echo "$tool conv rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_conv.log
done
