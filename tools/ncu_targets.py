#!/usr/bin/env python3
"""Short workloads to run under ncu (one GPU): `sam` launches the stand-alone soft-argmax kernel on
rotating config-B inputs, `net` runs two config-B inference steps, `conv` one block-3/4 sized layer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def sam(side=16, stride=16, j=17, n=256, dt='f32'):
    from metro_pose3d_b200.inference import SoftArgmax
    from metro_pose3d_b200.weights import synth_head
    base = torch.from_numpy(synth_head(8, side, j, seed=0)).cuda()
    heads = [base.repeat((n + 7) // 8, 1, 1, 1)[:n].roll(r, 0).contiguous() for r in range(6)]
    if dt == 'f16':
        heads = [h.half() for h in heads]
    op = SoftArgmax(side, j, stride, list(range(j)), head_dtype=dt)
    out = torch.empty((n, j, 3), device='cuda')
    for h in heads:
        op(h, out)
    torch.cuda.synchronize()


def net(arch='resnet_v2_50', stride=16, ds='h36m', n=256, steps=2):
    from metro_pose3d_b200.inference import MetroModel
    model = MetroModel(arch, stride, ds, max_batch=n)
    x = [torch.rand((n, 256, 256, 3), device='cuda') for _ in range(2)]
    for i in range(steps):
        model.infer(x[i & 1])
    torch.cuda.synchronize()


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'sam'
    if what == 'sam':
        sam()
    elif what == 'sam_e':
        sam(64, 4, 19, 128)
    elif what == 'net':
        net(n=int(sys.argv[2]) if len(sys.argv) > 2 else 256, steps=int(sys.argv[3]) if len(sys.argv) > 3 else 2)
    elif what == 'net_d':
        net('resnet_v2_101', 16, 'coco19')
    elif what == 'net_cfg':            # net_cfg <config> <batch> <steps>
        from metro_pose3d_b200.spec import CONFIGS
        arch, stride, ds, _, _ = CONFIGS[sys.argv[2]]
        net(arch, stride, ds, n=int(sys.argv[3]), steps=int(sys.argv[4]) if len(sys.argv) > 4 else 2)
    elif what == 'roles':              # METRO_ROLE_PROF=1: per-launch times + role timers of one config-B step
        from metro_pose3d_b200.inference import MetroModel
        m = MetroModel('resnet_v2_50', 16, 'h36m', max_batch=256)
        x = torch.rand((256, 256, 256, 3), device='cuda')
        m.infer(x); torch.cuda.synchronize()
        for name, ms in m.profile(x):
            print(f'{name:50s} {ms*1e3:9.1f} us')
