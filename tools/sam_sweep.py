#!/usr/bin/env python3
"""Stand-alone soft-argmax timing (CUDA-graph replay, rotating inputs > L2).  Sweeps the plan knobs the C-ABI
exposes (lanes, word_bytes, splits) for both head dtypes; ring parameters come from the environment
(METRO_SAM_STAGES, METRO_SAM_TILE_KB).  Prints us/launch, GB/s and the error against the oracle.
usage: sam_sweep.py [quick]"""
import itertools, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metro_pose3d_b200.inference import SoftArgmax
from metro_pose3d_b200.weights import synth_head
from oracle.metro_oracle import decode_ref

quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'
# optional 2nd argument: knob triples "lanes,word_bytes,splits;..." ; optional 3rd: comma list of METRO_SAM_CH values
tag = f"stages={os.environ.get('METRO_SAM_STAGES', '-')} tile_kb={os.environ.get('METRO_SAM_TILE_KB', '-')}"
shapes = [(16, 16, 17, 256)] if quick else [(16, 16, 17, 256), (16, 16, 19, 256), (32, 8, 19, 64), (64, 4, 19, 128)]
knobs = [(0, 0, 0)] + [(l, w, 0) for l, w in itertools.product((2, 4, 8), (8, 16))]
if len(sys.argv) > 2:
    knobs = [tuple(int(v) for v in k.split(',')) for k in sys.argv[2].split(';')]
chs = sys.argv[3].split(',') if len(sys.argv) > 3 else [os.environ.get('METRO_SAM_CH', '')]
for (side, stride, j, n), dt in itertools.product(shapes, ('f32', 'f16')):
    perm = list(range(j))
    base = synth_head(8, side, j, seed=0)
    want = decode_ref(base, j, stride, perm)
    tdt = torch.float32 if dt == 'f32' else torch.float16
    if dt == 'f16':
        want = decode_ref(base.astype(np.float16).astype(np.float32), j, stride, perm)
    nbytes = n * side * side * 8 * j * (4 if dt == 'f32' else 2)
    nrot = max(2, int(np.ceil(400e6 / nbytes)))
    tb = torch.from_numpy(base).cuda().to(tdt)
    heads = [tb.repeat((n + 7) // 8, 1, 1, 1)[:n].roll(r, 0).contiguous() for r in range(nrot)]
    out = torch.empty((n, j, 3), device='cuda')
    for (lanes, wb, splits), ch in itertools.product(knobs, chs):
        if ch:
            os.environ['METRO_SAM_CH'] = ch
        tag = f'ch={ch or "-"}'
        try:
            op = SoftArgmax(side, j, stride, perm, head_dtype=dt, lanes=lanes, word_bytes=wb, splits=splits)
            got = op(tb).cpu().numpy()
        except Exception as e:  # a knob combination the plan rejects
            print(f'{tag}: side={side} J={j} {dt} lanes={lanes} wb={wb}: rejected ({e})', flush=True)
            continue
        err = np.abs(got - want).max()
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
        print(f'{tag}: side={side} J={j} n={n} {dt} lanes={lanes} wb={wb} splits={splits}: {us:7.2f} us '
              f'{nbytes / us / 1e3:7.1f} GB/s  max err {err:.2e} mm', flush=True)
