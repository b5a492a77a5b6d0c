#!/usr/bin/env bash
# first GPU call of a session: parity tests, bench, launch list, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench rc=$?"
cat gpurun_out/bench_B.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_net.csv python tools/ncu_targets.py net > gpurun_out/ncu_net.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:softargmax -s 6 -c 3 -o gpurun_out/sam_full -f python tools/ncu_targets.py sam > gpurun_out/ncu_sam.log 2>&1; echo "ncu sam rc=$?"
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none -k regex:conv_gemm -s 54 -c 54 -o gpurun_out/net_sol -f python tools/ncu_targets.py net > gpurun_out/ncu_net_sol.log 2>&1; echo "ncu net rc=$?"
