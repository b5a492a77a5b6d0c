#!/usr/bin/env bash
# round-end evidence: full GPU test suite, bench lines (ours + reference arm), ncu launch list with DRAM bytes,
# full ncu captures of the representative kernels.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
for c in D C E; do timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "config $c rc=$?"; done
[ -n "$QUICK" ] || { timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_net.csv python tools/ncu_targets.py net 256 2 > /dev/null 2>&1; echo "ncu list rc=$?"
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/full_$1 -f python tools/ncu_targets.py $5 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu $1 rc=$?"; }
[ -n "$QUICK" ] || { cap root root_fused 1 1 net; cap b1u2 conv_gemm 52 3 net; cap b3u2 conv_gemm 73 3 net; cap b4u2 conv_gemm 91 3 net; }
cap sam softargmax 3 2 sam
timeout 300 python tools/sam_sweep.py all "0,0,0" > gpurun_out/sam_sweep_final.log 2>&1; echo "sam sweep rc=$?"
METRO_SAM_PROF=1 timeout 300 python tools/sam_prof.py > gpurun_out/sam_prof_final.log 2>&1; echo "sam prof rc=$?"
