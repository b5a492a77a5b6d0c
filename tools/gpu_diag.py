#!/usr/bin/env python3
"""Staged GPU diagnostics (each stage in its own process with a timeout, so a trapped kernel in one
stage cannot take the others down).  Writes gpurun_out/diag_<stage>.log."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'gpurun_out')


def stage_sam():
    import numpy as np, torch
    from metro_pose3d_b200.inference import SoftArgmax
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.weights import synth_head
    from oracle.metro_oracle import decode_ref
    for side, stride, j, ds, n in [(8, 32, 17, 'h36m', 5), (16, 16, 17, 'h36m', 7), (32, 8, 19, 'coco19', 4), (64, 4, 19, 'coco19', 2)]:
        perm = export_permutation(ds)
        x = synth_head(n, side, j, seed=side + j)
        for dt in ('f32', 'f16'):
            xx = x.astype(np.float16).astype(np.float32) if dt == 'f16' else x
            t = torch.from_numpy(xx).cuda()
            t = t.half() if dt == 'f16' else t
            got = SoftArgmax(side, j, stride, perm, head_dtype=dt)(t).cpu().numpy()
            ref = decode_ref(xx, j, stride, perm)
            print(f'sam side={side} J={j} {dt}: max err {np.abs(got-ref).max():.3e} mm', flush=True)
    # timing sweep, config B and D/E-like shapes, rotating inputs > L2; launches replayed from a CUDA graph
    # so the figure is device time per launch, not Python/ctypes call overhead
    for side, stride, j, n in [(16, 16, 17, 256), (16, 16, 19, 256), (32, 8, 19, 64), (64, 4, 19, 128)]:
        perm = list(range(j))
        base = torch.from_numpy(synth_head(8, side, j, seed=0)).cuda()
        for dt in ('f32', 'f16'):
            isz = 4 if dt == 'f32' else 2
            nbytes = n * side * side * 8 * j * isz
            nrot = max(2, int(np.ceil(400e6 / nbytes)))
            heads = []
            for r in range(nrot):
                h = base.repeat((n + 7) // 8, 1, 1, 1)[:n].roll(r, 0).contiguous()
                heads.append(h.half() if dt == 'f16' else h)
            P = side * side
            for lanes, splits, wb in [(0, 0, 0), (2, 0, 8), (4, 0, 8), (2, 0, 16), (4, 1, 16), (4, 2, 16), (4, 4, 16), (4, 8, 16), (2, 2, 8), (2, 4, 8)]:
                try:
                    op = SoftArgmax(side, j, stride, perm, head_dtype=dt, splits=splits, lanes=lanes, word_bytes=wb)
                    out = torch.empty((n, j, 3), device='cuda')
                    for r in range(nrot): op(heads[r], out)
                    torch.cuda.synchronize()
                    it = 4 * nrot
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for i in range(it): op(heads[i % nrot], out)
                    g.replay(); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(5): g.replay()
                    e1.record(); torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) / (5 * it) * 1e3
                    print(f'sam-time side={side} J={j} n={n} {dt} lanes={lanes} splits={splits} wb={wb}: {us:8.2f} us  {nbytes/us/1e3:8.1f} GB/s', flush=True)
                except Exception as e:
                    print(f'sam-time side={side} J={j} {dt} lanes={lanes} splits={splits}: ERROR {e}', flush=True)


def _conv_case(n, side, cin, cout, k, stride, rate, pad_lo, verbose=True):
    import numpy as np, torch
    from metro_pose3d_b200.inference import conv2d
    from oracle.metro_oracle import conv2d_fused_ref
    rng = np.random.default_rng(cin + cout + k)
    x = rng.standard_normal((n, side, side, cin)).astype(np.float16)
    w = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    k_eff = k + (k - 1) * (rate - 1)
    lo = (k_eff - 1) // 2 if pad_lo is None else pad_lo
    hi = (k_eff - 1) - lo
    y = conv2d(torch.from_numpy(x).cuda(), w, scale, shift, stride=stride, rate=rate, pad_lo=lo, relu=False)
    torch.cuda.synchronize()
    y = y.float().cpu().numpy()
    ref, _ = conv2d_fused_ref(x, w, scale, shift, stride, rate, lo, hi)
    err = np.abs(y - ref)
    tol = np.abs(ref) * 2.0 ** -10 + 1e-5
    nbad = int((err > tol).sum())
    print(f'conv n={n} side={side} cin={cin} cout={cout} k={k} s={stride} r={rate} pad={lo}: max err {err.max():.3e} '
          f'bad {nbad}/{err.size} ref rms {np.sqrt((ref**2).mean()):.3f}', flush=True)
    if nbad and verbose:
        idx = np.argwhere(err > tol)
        print('   first bad idx', idx[:5].tolist(), 'got', y[tuple(idx[0])], 'ref', ref[tuple(idx[0])])
        print('   bad per image', [(int((err[i] > tol).sum())) for i in range(n)])
        print('   bad rows(h) hist', np.bincount(idx[:, 1], minlength=y.shape[1]).tolist())
        print('   bad cols(w) hist', np.bincount(idx[:, 2], minlength=y.shape[2]).tolist())
        print('   bad chan hist/8', np.bincount(idx[:, 3] // 8, minlength=(cout + 7) // 8).tolist())
        print('   y[0,0,0,:8]', y[0, 0, 0, :8], '\n   r[0,0,0,:8]', ref[0, 0, 0, :8])
        # is y a permuted / scaled version?  correlation of the whole tensor
        print('   corr(y, ref) =', float(np.corrcoef(y.ravel(), ref.ravel())[0, 1]))
    return nbad


def stage_conv1x1():
    for c in [(2, 16, 64, 64, 1, 1, 1, None), (2, 16, 128, 128, 1, 1, 1, None), (3, 8, 128, 256, 1, 1, 1, None),
              (2, 16, 512, 2048, 1, 1, 1, None), (2, 16, 2048, 136, 1, 1, 1, None)]:
        _conv_case(*c)


def stage_conv3x3():
    for c in [(2, 64, 64, 64, 3, 1, 1, None), (2, 16, 128, 128, 3, 1, 2, None), (1, 64, 64, 64, 3, 1, 8, None)]:
        _conv_case(*c)


def stage_convs2():
    for c in [(2, 32, 128, 128, 3, 2, 1, 1), (2, 32, 128, 128, 3, 2, 1, 0), (3, 16, 256, 256, 3, 2, 1, 0)]:
        _conv_case(*c)


def stage_net():
    import numpy as np, torch
    from metro_pose3d_b200.inference import MetroModel
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.spec import NetSpec
    from metro_pose3d_b200.weights import synth_images, synth_weights
    from oracle.metro_oracle import OracleNet
    for arch, stride in [('resnet_v2_50', 32), ('resnet_v2_50', 16)]:
        spec = NetSpec(arch, stride, 17)
        w = synth_weights(spec, 0)
        img = synth_images(2, seed=1000)
        model = MetroModel(arch, stride, 'h36m', weights=w, max_batch=2, keep_activations=True)
        poses = model.infer(torch.from_numpy(img).cuda())
        torch.cuda.synchronize()
        poses = poses.cpu().numpy()
        ora = OracleNet(spec, w, export_permutation('h36m'), 'half')
        ora.trace = {}
        head = ora.forward_head(img)
        ref = ora.decode(head)
        for name, t in ora.trace.items():
            if name == 'postnorm':
                continue
            got = model.debug_read(name).reshape(t.shape).astype(np.float64)
            rel = np.linalg.norm(got - t) / max(np.linalg.norm(t), 1e-30)
            print(f'net {arch} s{stride} {name:24s} rel err {rel:.3e} max abs {np.abs(got-t).max():.3e}', flush=True)
        gh = model.debug_read('head').reshape(head.shape)
        print(f'net head rel err {np.linalg.norm(gh-head)/np.linalg.norm(head):.3e}')
        print(f'net {arch} s{stride} poses max err vs half-oracle {np.abs(poses-ref).max():.4f} mm', flush=True)
        p64 = OracleNet(spec, w, export_permutation('h36m'), 'fp64')(img)
        print(f'net {arch} s{stride} poses max err vs fp64-oracle {np.abs(poses-p64).max():.4f} mm ; half-oracle vs fp64 {np.abs(ref-p64).max():.4f} mm', flush=True)
        model.close()


def stage_prof():
    import numpy as np, torch
    from metro_pose3d_b200.inference import MetroModel
    from metro_pose3d_b200.spec import NetSpec
    for arch, stride, ds, n in [('resnet_v2_50', 16, 'h36m', 256), ('resnet_v2_101', 16, 'coco19', 256), ('resnet_v2_50', 8, 'coco19', 64)]:
        j = 17 if ds == 'h36m' else 19
        spec = NetSpec(arch, stride, j)
        model = MetroModel(arch, stride, ds, max_batch=n)
        x = torch.rand((n, 256, 256, 3), device='cuda')
        out = model.infer(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): model.infer(x, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'prof {arch} s{stride} n={n}: {ms:.3f} ms/step  {n/ms*1e3:.0f} crops/s  {spec.flops_per_crop*n/ms/1e9:.1f} TFLOP/s', flush=True)
        model.profile(x)
        acc = {}
        for _ in range(3):
            for name, t in model.profile(x):
                acc[name] = acc.get(name, 0) + t / 3
        flops = {c.name: c.flops for c in spec.convs}
        for u in spec.units:
            if u.shortcut is not None:
                flops[u.conv3.name] += flops[u.shortcut.name]
        for name, t in acc.items():
            f = flops.get(name, 0) * n
            print(f'   {name:28s} {t*1e3:9.1f} us  {f/t/1e9 if t > 0 else 0:8.1f} TFLOP/s', flush=True)
        model.close()


STAGES = {'sam': stage_sam, 'conv1x1': stage_conv1x1, 'conv3x3': stage_conv3x3, 'convs2': stage_convs2,
          'net': stage_net, 'prof': stage_prof}

if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--stage':
        STAGES[sys.argv[2]]()
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(STAGES)
    for s in names:
        t0 = time.time()
        log = os.path.join(OUT, f'diag_{s}.log')
        with open(log, 'w') as f:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), '--stage', s], stdout=f,
                                   stderr=subprocess.STDOUT, timeout=420)
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = 'TIMEOUT'
        print(f'=== stage {s}: rc={rc} ({time.time()-t0:.0f}s)')
        print(open(log).read()[-6000:])
