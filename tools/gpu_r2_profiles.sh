#!/usr/bin/env bash
# round 2 evidence run: full GPU suite, bench lines (ours + reference arm + other configs), ncu launch list with DRAM bytes,
# full ncu captures of the representative kernels, compute-sanitizer passes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.json
timeout 1200 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
[ -n "$QUICK" ] || { timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; }
for c in D C E A; do timeout 600 python bench.py --config $c --steps 5 --no-cpu-baseline --layers > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "config $c rc=$?"; done
for c in B D C E; do
  n=256; [ $c = C ] && n=64; [ $c = E ] && n=128
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_net_$c.csv python tools/ncu_targets.py net_cfg $c $n 2 > /dev/null 2>&1; echo "ncu list $c rc=$?"
done
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/full_$1 -f python tools/ncu_targets.py $5 > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu $1 rc=$?"; }
if [ -z "$QUICK" ]; then
  cap root root_fused 1 1 net
  cap chain conv_chain 9 3 net        # second step: block1 u2->u3, block2 u1->u2, u2->u3
  cap chain3 conv_chain 13 2 net      # block3 u1->u2, u2->u3
  cap b3c2 conv_gemm 62 2 net         # around block3 conv2 of the second step
  cap sam softargmax 3 2 sam
fi
timeout 300 python tools/sam_sweep.py all "0,0,0" > gpurun_out/sam_sweep_final.log 2>&1; echo "sam sweep rc=$?"
METRO_SAM_PROF=1 timeout 300 python tools/sam_prof.py > gpurun_out/sam_prof_final.log 2>&1; echo "sam prof rc=$?"
METRO_ROLE_PROF=1 timeout 300 python tools/ncu_targets.py roles > gpurun_out/roles.log 2> gpurun_out/roles.txt; echo "roles rc=$?"
# compute-sanitizer: every kernel family incl. the chained / strict / crop / dataflow paths
CS="compute-sanitizer --print-limit 10"
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool python tools/ncu_targets.py net 3 1 > gpurun_out/sanitizer_${tool}_net.log 2>&1; echo "$tool net rc=$? $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_net.log | tail -1)"
  timeout 600 $CS --tool $tool python -m pytest tests/test_softargmax_gpu.py tests/test_crops.py tests/test_post.py -m gpu -q -x -k "cross_cta or known or parity_vs or crops_bit or back_project or coords_and" > gpurun_out/sanitizer_${tool}_ops.log 2>&1; echo "$tool ops rc=$? $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_ops.log | tail -1)"
done
timeout 600 $CS --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "layer_by_layer" > gpurun_out/sanitizer_memcheck_strict.log 2>&1; echo "memcheck strict rc=$? $(grep -h 'ERROR SUMMARY' gpurun_out/sanitizer_memcheck_strict.log | tail -1)"
