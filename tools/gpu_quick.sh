#!/usr/bin/env bash
# quick check: conv + net parity tests, then the bench line with per-layer times
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_net_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_B.json').read())
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'conv frac', round(d['roofline']['frac'], 3), 'sam', round(d['roofline_softargmax']['achieved']), d['clocks'])
PY
grep " us$" gpurun_out/bench_B.err | awk '{printf "%-26s %8.1f   ", $1, $2; if (NR%3==0) printf "\n"} END{printf "\n"}'
