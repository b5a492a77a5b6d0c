#!/usr/bin/env bash
# role-timer profile of one config-B step (METRO_ROLE_PROF=1) -> gpurun_out/roles.txt
mkdir -p gpurun_out
METRO_ROLE_PROF=1 timeout 300 python - > gpurun_out/roles.log 2> gpurun_out/roles.txt <<'PY'
import sys, torch
sys.path.insert(0, '.')
from metro_pose3d_b200.inference import MetroModel
m = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=256)
x = torch.rand((256, 256, 256, 3), device='cuda')
m.infer(x); torch.cuda.synchronize()
for name, ms in m.profile(x):
    print(f'{name:28s} {ms*1e3:9.1f} us')
PY
echo "rc=$?"; head -52 gpurun_out/roles.txt
