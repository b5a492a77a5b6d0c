"""GPU parity: fused soft-argmax kernel (through the C-ABI metro_softargmax) vs the float64 oracle.
Tolerance: 1e-3 mm per coordinate (BASELINE.json north_star)."""
import numpy as np
import pytest

from metro_pose3d_b200.joints import export_permutation
from metro_pose3d_b200.weights import synth_head
from oracle.metro_oracle import decode_ref

pytestmark = pytest.mark.gpu
TOL_MM = 1e-3


def _run(x, j, stride, perm, **kw):
    import torch
    from metro_pose3d_b200.inference import SoftArgmax
    dt = kw.get('head_dtype', 'f32')
    t = torch.from_numpy(x).cuda()
    if dt == 'f16':
        t = t.half()
    op = SoftArgmax(x.shape[1], j, stride, perm, **kw)
    out = op(t)
    torch.cuda.synchronize()
    return out.cpu().numpy()


CASES = [  # side, stride, J_model, dataset, n   (BASELINE configs A..E heads + merged 53->19)
    (8, 32, 17, 'h36m', 5), (16, 16, 17, 'h36m', 7), (32, 8, 19, 'coco19', 4), (16, 16, 19, 'coco19', 3),
    (64, 4, 19, 'coco19', 2), (8, 32, 53, 'merged', 3),
]


@pytest.mark.parametrize('side,stride,j,ds,n', CASES)
def test_parity_vs_float64_oracle(side, stride, j, ds, n):
    perm = export_permutation(ds)
    x = synth_head(n, side, j, seed=side + j)
    got = _run(x, j, stride, perm)
    ref = decode_ref(x, j, stride, perm)
    err = np.abs(got - ref).max()
    assert err < TOL_MM, f'max |err| = {err:.3e} mm'


@pytest.mark.parametrize('side,stride,j,ds,n', CASES[:5])
def test_parity_fp16_head(side, stride, j, ds, n):
    perm = export_permutation(ds)
    x = synth_head(n, side, j, seed=1 + side).astype(np.float16).astype(np.float32)
    got = _run(x, j, stride, perm, head_dtype='f16')
    ref = decode_ref(x, j, stride, perm)
    assert np.abs(got - ref).max() < TOL_MM


@pytest.mark.parametrize('splits,lanes', [(1, 11), (2, 8), (5, 4), (16, 1), (3, 0), (8, 8), (4, 16)])
def test_cross_cta_merge_paths(splits, lanes):
    """Different (splits, lanes) exercise the single-CTA path, the cluster merge through distributed shared memory
    (2, 4, 8 CTAs per crop), the ticketed multi-CTA merge of the other split counts, the wide-CTA variant and ragged
    tails."""
    perm = export_permutation('h36m')
    x = synth_head(6, 16, 17, seed=11, sigma=4.0)
    ref = decode_ref(x, 17, 16, perm)
    got = _run(x, 17, 16, perm, splits=splits, lanes=lanes)
    assert np.abs(got - ref).max() < TOL_MM


def test_known_answers():
    perm = export_permutation('h36m')
    x = np.full((2, 16, 16, 136), -3.25, np.float32)
    assert np.abs(_run(x, 17, 16, perm)).max() < 1e-4            # uniform -> all coords 0.5 -> zeros
    x = np.zeros((1, 16, 16, 136), np.float32)
    peaks = {}
    rng = np.random.default_rng(0)
    for jj in range(17):
        h, w, d = (int(v) for v in (rng.integers(16), rng.integers(16), rng.integers(8)))
        peaks[jj] = (h, w, d)
        x[0, h, w, d * 17 + jj] = 1e4
    got = _run(x, 17, 16, perm)
    sx = 239 * 2200.0 / 256
    for jo, jm in enumerate(perm):
        h, w, d = peaks[jm]
        hr, wr, dr = peaks[16]
        want = [(w - wr) / 15 * sx, (h - hr) / 15 * sx, (d - dr) / 7 * 2200.0]
        assert np.abs(got[0, jo] - want).max() < TOL_MM
    big = _run(synth_head(2, 16, 17, seed=2) * 300.0, 17, 16, perm)   # |logits| ~ 1e3: no overflow
    assert np.isfinite(big).all()


def test_full_size_properties_config_B():
    """N=256, 16x16x136 (BASELINE config B): results are independent of batch position / sharding and
    repeated launches reuse the self-cleaning workspace."""
    import torch
    from metro_pose3d_b200.inference import SoftArgmax
    perm = export_permutation('h36m')
    x = torch.from_numpy(synth_head(256, 16, 17, seed=0)).cuda()
    op = SoftArgmax(16, 17, 16, perm)
    a = op(x).clone()
    b = op(x).clone()
    assert torch.equal(a, b)
    p = torch.randperm(256, device='cuda')
    assert torch.equal(op(x[p].contiguous()), a[p])
    halves = torch.cat([SoftArgmax(16, 17, 16, perm)(x[:128].contiguous()),
                        SoftArgmax(16, 17, 16, perm)(x[128:].contiguous())])
    assert torch.equal(halves, a)
    ref = decode_ref(x[:16].cpu().numpy(), 17, 16, perm)
    assert np.abs(a[:16].cpu().numpy() - ref).max() < TOL_MM
    assert torch.all(a[:, 0] == 0)


def test_large_heatmap_config_E_slice():
    perm = export_permutation('coco19')
    x = synth_head(8, 64, 19, seed=4)
    got = _run(x, 19, 4, perm)
    ref = decode_ref(x, 19, 4, perm)
    assert np.abs(got - ref).max() < TOL_MM


def test_argument_errors():
    import torch
    from metro_pose3d_b200.inference import SoftArgmax
    with pytest.raises(ValueError):
        SoftArgmax(16, 17, 16, [0, 99])(torch.zeros(1, 16, 16, 136, device='cuda'))
    op = SoftArgmax(16, 17, 16, [0])
    with pytest.raises(ValueError):
        op(torch.zeros(1, 8, 8, 136, device='cuda'))
    assert op(torch.zeros(0, 16, 16, 136, device='cuda')).shape == (0, 1, 3)   # empty batch
