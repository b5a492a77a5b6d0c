"""GPU parity: the fused tcgen05 implicit-GEMM convolution (C-ABI metro_conv2d) vs the float64
operator oracle evaluated on the same fp16 operands.  Tolerance: the fp32 tensor-core accumulation
must land within one fp16 rounding step of the exact value (|err| <= 2^-10 |y| + 2^-10 * 1e-2), or
2e-6 relative for float32 outputs."""
import numpy as np
import pytest

from oracle.metro_oracle import conv2d_fused_ref

pytestmark = pytest.mark.gpu


def _tol16(ref):
    return np.abs(ref) * 2.0 ** -10 + 1e-5


def _mk(n, side, cin, cout, k, seed, cin2=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, side, side, cin)).astype(np.float16)
    w = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = (0.1 * rng.standard_normal(cout)).astype(np.float32)
    return rng, x, w, scale, shift


def _gpu(x):
    import torch
    return torch.from_numpy(x).cuda()


CASES = [
    # n, side, cin, cout, k, stride, rate, pad_lo(None=SAME)
    (2, 16, 64, 64, 1, 1, 1, None),      # smallest GEMM: one K block, one tile per 128 pixels
    (3, 8, 128, 256, 1, 1, 1, None),     # 8x8 maps: two crops per tile, odd batch -> ragged last tile
    (1, 32, 256, 128, 1, 1, 1, None),
    (2, 16, 512, 2048, 1, 1, 1, None),   # many N tiles
    (2, 64, 64, 64, 3, 1, 1, None),      # block1 conv2
    (2, 16, 128, 128, 3, 1, 2, None),    # atrous rate 2
    (1, 32, 64, 64, 3, 1, 4, None),      # atrous rate 4
    (1, 64, 64, 64, 3, 1, 8, None),      # atrous rate 8 (config E block4)
    (2, 32, 128, 128, 3, 2, 1, 1),       # stride 2, explicit pad (1,1)  (Q4)
    (2, 32, 128, 128, 3, 2, 1, 0),       # stride 2, centred: TF SAME pad (0,1)  (Q5)
    (3, 16, 256, 256, 3, 2, 1, 0),       # 16 -> 8: two crops per tile
    (2, 16, 2048, 136, 1, 1, 1, None),   # logits head, J=17 (BLOCK_N 160, masked columns)
    (2, 16, 2048, 152, 1, 1, 1, None),   # logits head, J=19
]


@pytest.mark.parametrize('n,side,cin,cout,k,stride,rate,pad_lo', CASES)
def test_conv_bn_relu(n, side, cin, cout, k, stride, rate, pad_lo):
    from metro_pose3d_b200.inference import conv2d
    _, x, w, scale, shift = _mk(n, side, cin, cout, k, seed=cin + cout + k + stride + rate)
    k_eff = k + (k - 1) * (rate - 1)
    lo = (k_eff - 1) // 2 if pad_lo is None else pad_lo
    hi = (k_eff - 1) - lo
    y = conv2d(_gpu(x), w, scale, shift, stride=stride, rate=rate, pad_lo=lo, relu=True).float().cpu().numpy()
    ref, _ = conv2d_fused_ref(x, w, scale, shift, stride, rate, lo, hi, relu=True)
    assert y.shape == ref.shape
    bad = np.abs(y - ref) > _tol16(ref)
    assert not bad.any(), f'{bad.sum()} / {bad.size} outside tolerance, max err {np.abs(y - ref).max():.3e}'


def test_float32_output_head():
    from metro_pose3d_b200.inference import conv2d
    _, x, w, _, shift = _mk(2, 16, 2048, 136, 1, seed=5)
    scale = np.ones(136, np.float32)
    y = conv2d(_gpu(x), w, scale, shift, out_dtype='f32').cpu().numpy()
    ref, _ = conv2d_fused_ref(x, w, scale, shift, out_f16=False)
    assert np.abs(y - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('res_stride,res_shift', [(1, 0), (2, 0), (2, 1)])
def test_residual_and_second_output(res_stride, res_shift):
    """conv3 + bias + identity shortcut (sub-sampled, centred offset) and the fused next pre-activation."""
    from metro_pose3d_b200.inference import conv2d
    rng, x, w, _, shift = _mk(2, 16, 64, 256, 1, seed=7 + res_stride + res_shift)
    scale = np.ones(256, np.float32)
    res = rng.standard_normal((2, 16 * res_stride, 16 * res_stride, 256)).astype(np.float16)
    s2 = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    f2 = (0.1 * rng.standard_normal(256)).astype(np.float32)
    y, y2 = conv2d(_gpu(x), w, scale, shift, res=_gpu(res), res_stride=res_stride, res_shift=res_shift,
                   scale2=s2, shift2=f2)
    ref, _ = conv2d_fused_ref(x, w, scale, shift, res_nhwc=res, res_stride=res_stride, res_shift=res_shift)
    y = y.float().cpu().numpy()
    assert not (np.abs(y - ref) > _tol16(ref)).any(), np.abs(y - ref).max()
    # second output is defined on the fp16 value actually stored
    ref2 = np.maximum(y.astype(np.float64) * s2 + f2, 0)
    assert not (np.abs(y2.float().cpu().numpy() - ref2) > _tol16(ref2)).any()


def test_projection_shortcut_as_second_source():
    """conv3(r2) + shortcut1x1(preact) + both biases in one accumulator (resnet_v2.py:123-125,134-138)."""
    from metro_pose3d_b200.inference import conv2d
    rng, x, w, _, shift = _mk(2, 32, 64, 256, 1, seed=21)
    x2 = rng.standard_normal((2, 32, 32, 128)).astype(np.float16)
    w2 = (rng.standard_normal((1, 1, 128, 256)) * 0.1).astype(np.float32)
    scale = np.ones(256, np.float32)
    y = conv2d(_gpu(x), w, scale, shift, x2=_gpu(x2), w2=w2).float().cpu().numpy()
    ref, _ = conv2d_fused_ref(x, w, scale, shift, x2_nhwc=x2, w2=w2)
    assert not (np.abs(y - ref) > _tol16(ref)).any(), np.abs(y - ref).max()


def test_full_size_layer_against_torch_fp32_and_linearity():
    """Config-B sized layer (M = 65536 pixels, K = 4608, N = 512; every CTA walks several tiles, so the
    TMEM double-buffering, the shared-memory ring wrap-around and the output staging reuse are exercised):
      * the whole batch against a plain PyTorch fp32 convolution of the same fp16 operands,
      * bit-exact repeatability, and batch-slice consistency (a crop's result does not depend on its tile),
      * linearity conv(2x) == 2 conv(x): exact wherever the result is a normal fp16 number (doubling moves
        a subnormal result onto a finer grid, so |y| < 2^-13 is excluded)."""
    import torch
    import torch.nn.functional as F
    from metro_pose3d_b200.inference import conv2d
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((3, 3, 512, 512)) * 0.02).astype(np.float32)
    one, zero = np.ones(512, np.float32), np.zeros(512, np.float32)
    x = (torch.randn(256, 16, 16, 512, device='cuda') * 0.5).half()
    y = conv2d(x, w, one, zero, rate=2)
    assert torch.equal(conv2d(x, w, one, zero, rate=2), y)
    y2 = conv2d((x * 2).half(), w, one, zero, rate=2)
    normal = y.float().abs() >= 2.0 ** -13
    assert torch.equal(y2.float()[normal], (y.float() * 2)[normal])
    ys = conv2d(x[100:104].contiguous(), w, one, zero, rate=2)
    assert torch.equal(ys, y[100:104])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    wt = torch.from_numpy(w).cuda().half().float().permute(3, 2, 0, 1).contiguous()
    for lo in range(0, 256, 64):
        ref = F.conv2d(x[lo:lo + 64].float().permute(0, 3, 1, 2), wt, padding=2, dilation=2).permute(0, 2, 3, 1)
        err = (y[lo:lo + 64].float() - ref).abs()
        assert bool((err <= ref.abs() * 2.0 ** -10 + 1e-3).all()), float(err.max())
    ref, _ = conv2d_fused_ref(x[:1].cpu().numpy(), w, one, zero, rate=2)
    got = y[:1].float().cpu().numpy()
    assert not (np.abs(got - ref) > _tol16(ref)).any()


def test_full_size_residual_unit_tail_against_torch_fp32():
    """conv3 + bias + identity shortcut + fused next pre-activation at block-3 size for the whole batch
    (M = 65536, N = 1024): many tiles per CTA with two TMA-stored outputs and the identity K blocks."""
    import torch
    from metro_pose3d_b200.inference import conv2d
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((1, 1, 256, 1024)) * 0.05).astype(np.float32)
    bias = (0.1 * rng.standard_normal(1024)).astype(np.float32)
    s2 = rng.uniform(0.5, 1.5, 1024).astype(np.float32)
    f2 = (0.1 * rng.standard_normal(1024)).astype(np.float32)
    x = torch.randn(256, 16, 16, 256, device='cuda').half()
    res = torch.randn(256, 16, 16, 1024, device='cuda').half()
    y, y2 = conv2d(x, w, np.ones(1024, np.float32), bias, res=res, res_stride=1, scale2=s2, shift2=f2)
    wt = torch.from_numpy(w).cuda().half().float().reshape(256, 1024)
    ref = x.float().reshape(-1, 256) @ wt + torch.from_numpy(bias).cuda() + res.float().reshape(-1, 1024)
    err = (y.float().reshape(-1, 1024) - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -10 + 1e-3).all()), float(err.max())
    ref2 = torch.relu(y.float() * torch.from_numpy(s2).cuda() + torch.from_numpy(f2).cuda())
    err2 = (y2.float() - ref2).abs()
    assert bool((err2 <= ref2.abs() * 2.0 ** -10 + 1e-3).all()), float(err2.max())
