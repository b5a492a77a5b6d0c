#!/usr/bin/env python3
"""ms/step of metro_infer (device buffers) for one config: `time_step.py [config] [batch] [steps]`.  Environment knobs
(METRO_STEM_CHUNK, METRO_NO_*) are read by the library at metro_create, so A/B runs are separate processes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metro_pose3d_b200.inference import MetroModel
from metro_pose3d_b200.spec import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else 'B'
arch, stride, ds, batch, gpus = CONFIGS[cfg]
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else (batch // gpus if gpus > 1 else batch)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
m = MetroModel(arch, stride, ds, max_batch=n)
x = [torch.rand((n, 256, 256, 3), device='cuda') for _ in range(2)]
out = torch.empty((n, m.n_joints_out, 3), device='cuda')
for i in range(5):
    m.infer(x[i & 1], out=out)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m.infer(x[i & 1], out=out)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('METRO_'))
print(f'config {cfg} n={n} [{tag}]: {best:.3f} ms/step  {n / best * 1e3:.0f} crops/s  checksum {float(out.abs().sum()):.6e}')
