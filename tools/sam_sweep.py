#!/usr/bin/env python3
"""Stand-alone soft-argmax timing (CUDA-graph replay, rotating inputs > L2) for the ring parameters given in the
environment (METRO_SAM_STAGES, METRO_SAM_TILE_KB); prints us/launch and GB/s for the BASELINE shapes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metro_pose3d_b200.inference import SoftArgmax
from metro_pose3d_b200.weights import synth_head
from oracle.metro_oracle import decode_ref

tag = f"stages={os.environ.get('METRO_SAM_STAGES', '2')} tile_kb={os.environ.get('METRO_SAM_TILE_KB', '40')}"
for side, stride, j, n in [(16, 16, 17, 256), (16, 16, 19, 256), (32, 8, 19, 64), (64, 4, 19, 128)]:
    perm = list(range(j))
    base = synth_head(8, side, j, seed=0)
    op = SoftArgmax(side, j, stride, perm)
    got = op(torch.from_numpy(base).cuda()).cpu().numpy()
    err = np.abs(got - decode_ref(base, j, stride, perm)).max()
    nbytes = n * side * side * 8 * j * 4
    nrot = max(2, int(np.ceil(400e6 / nbytes)))
    tb = torch.from_numpy(base).cuda()
    heads = [tb.repeat((n + 7) // 8, 1, 1, 1)[:n].roll(r, 0).contiguous() for r in range(nrot)]
    out = torch.empty((n, j, 3), device='cuda')
    for h in heads:
        op(h, out)
    torch.cuda.synchronize()
    it = 4 * nrot
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(it):
                op(heads[i % nrot], out)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            g.replay()
        e1.record(st)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * it) * 1e3
    print(f'{tag}: side={side} J={j} n={n}: {us:7.2f} us {nbytes / us / 1e3:7.1f} GB/s  max err {err:.2e} mm', flush=True)
