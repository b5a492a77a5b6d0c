#!/usr/bin/env bash
# full ncu captures of representative layers + small-batch (L2-resident) launch lists
mkdir -p gpurun_out
cap() {  # name skip count
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s $2 -c $3 -o gpurun_out/full_$1 -f python tools/ncu_targets.py net > gpurun_out/ncu_full_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap root 50 1
cap b1u2 54 3
cap b3u2 75 3
cap b4u2 93 3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:softargmax -s 3 -c 2 -o gpurun_out/full_sam -f python tools/ncu_targets.py sam > gpurun_out/ncu_full_sam.log 2>&1; echo "ncu sam rc=$?"
for n in 8 16 32 64; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/launches_net_n$n.csv python tools/ncu_targets.py net $n 3 > gpurun_out/ncu_net_n$n.log 2>&1; echo "ncu list n=$n rc=$?"
done
