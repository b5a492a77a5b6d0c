"""Per-CTA phase timers of the soft-argmax (run with METRO_SAM_PROF=1): stream / records / merge / output cycles and
the first-CTA-start to last-CTA-end span, for the BASELINE shapes and two split settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from metro_pose3d_b200.inference import SoftArgmax
from metro_pose3d_b200.weights import synth_head
for side, stride, j, n, dt, sp in [(16,16,17,256,'f32',0),(16,16,17,256,'f16',0),(16,16,17,256,'f32',4),(64,4,19,128,'f32',2),(64,4,19,128,'f32',16)]:
    base = synth_head(8, side, j, seed=0)
    tb = torch.from_numpy(base).cuda().to(torch.float32 if dt=='f32' else torch.float16)
    h = tb.repeat(n // 8, 1, 1, 1).contiguous()
    op = SoftArgmax(side, j, stride, list(range(j)), head_dtype=dt, splits=sp)
    print(side, j, n, dt, 'splits', sp, flush=True)
    for _ in range(3):
        op(h); torch.cuda.synchronize()
