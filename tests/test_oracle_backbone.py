"""Pins the torch-based backbone oracle to the documented TensorFlow semantics the reference relies on
(SURVEY.md 8c): conv2d_same equivalence (resnet_utils.py:90-106), centred vs non-centred stride (Q4/Q5),
zero-padded max-pool (Q6), and an independent scalar-loop convolution."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from metro_pose3d_b200.joints import export_permutation
from metro_pose3d_b200.spec import NetSpec
from metro_pose3d_b200.weights import synth_weights, synth_images, pack_blob, unpack_blob
from oracle.metro_oracle import OracleNet, conv2d_fused_ref, conv2d_naive


@pytest.mark.parametrize('k,stride,rate,pad', [(1, 1, 1, (0, 0)), (3, 1, 1, (1, 1)), (3, 1, 2, (2, 2)),
                                               (3, 2, 1, (1, 1)), (3, 2, 1, (0, 1)), (7, 2, 1, (3, 3))])
def test_torch_conv_matches_scalar_loop(k, stride, rate, pad):
    rng = np.random.default_rng(k * 100 + stride * 10 + rate)
    x = rng.standard_normal((2, 10, 10, 5))
    w = rng.standard_normal((k, k, 5, 7))
    ref = conv2d_naive(x, w, stride, rate, pad[0], pad[1])
    xt = torch.from_numpy(x).permute(0, 3, 1, 2)
    wt = torch.from_numpy(w).permute(3, 2, 0, 1).contiguous()
    got = F.conv2d(F.pad(xt, (pad[0], pad[1], pad[0], pad[1])), wt, stride=stride, dilation=rate)
    assert np.abs(got.permute(0, 2, 3, 1).numpy() - ref).max() < 1e-10


def test_conv2d_same_equals_same_conv_then_subsample():
    """resnet_utils.py:90-106: conv2d_same(x, k, stride) == conv2d(x, k, 1, 'SAME')[::stride]."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 12, 12, 4))
    w = rng.standard_normal((3, 3, 4, 6))
    dense = conv2d_naive(x, w, 1, 1, 1, 1)
    strided = conv2d_naive(x, w, 2, 1, 1, 1)           # explicit pad (1,1) + VALID, stride 2 (Q4)
    assert np.allclose(strided, dense[:, ::2, ::2])


def test_centered_stride_is_a_one_pixel_window_shift():
    """Q5: TF SAME with stride 2 on an even input pads (0,1): windows centred on odd pixels."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 12, 12, 4))
    w = rng.standard_normal((3, 3, 4, 6))
    dense = conv2d_naive(x, w, 1, 1, 1, 1)
    centred = conv2d_naive(x, w, 2, 1, 0, 1)
    assert np.allclose(centred, dense[:, 1::2, 1::2])


def test_zero_padded_pool_clamps_borders():
    """Q6: pool1 pads with zeros, so an all-negative input gives 0 on the top/left border only."""
    sp = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(sp, 0)
    net = OracleNet(sp, w, export_permutation('h36m'), 'fp64')
    x = -torch.ones(1, 64, 128, 128, dtype=torch.float64)
    y = F.max_pool2d(F.pad(x, (1, 1, 1, 1)), 3, 2)
    assert y.shape[-1] == 64
    assert (y[..., 0, :] == 0).all() and (y[..., :, 0] == 0).all()
    assert (y[..., 1:, 1:] == -1).all()                 # pad_hi row/col is never reached: (128+2-3)//2+1 = 64


def test_oracle_modes_agree_and_trace_shapes():
    sp = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(sp, 0)
    img = synth_images(1)
    perm = export_permutation('h36m')
    n64 = OracleNet(sp, w, perm, 'fp64')
    n64.trace = {}
    p64 = n64(img)
    p32 = OracleNet(sp, w, perm, 'fp32')(img)
    p16 = OracleNet(sp, w, perm, 'half')(img)
    assert p64.shape == (1, 17, 3) and np.all(p64[:, 0] == 0)
    assert np.abs(p64 - p32).max() < 5e-3                # fp32 graph ~ 1e-3 mm from exact
    assert np.abs(p64 - p16).max() < 10.0                # fp16 storage (the reference default) ~ 1 mm
    assert n64.trace['conv1'].shape == (1, 128, 128, 64)
    assert n64.trace['pool1'].shape == (1, 64, 64, 64)
    assert n64.trace['block3/unit_6/out'].shape == (1, 8, 8, 1024)
    assert n64.trace['postnorm'].shape == (1, 8, 8, 2048)


def test_blob_roundtrip():
    sp = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(sp, 3)
    w2 = unpack_blob(sp, pack_blob(sp, w))
    assert all(np.array_equal(w[k], w2[k]) for k in w)


def test_fused_conv_ref_matches_unit_composition():
    """conv2d_fused_ref (operator oracle) composes to the same thing OracleNet._unit computes."""
    rng = np.random.default_rng(4)
    x = rng.standard_normal((1, 8, 8, 64)).astype(np.float16)
    w = (rng.standard_normal((3, 3, 64, 64)) * 0.05).astype(np.float32)
    y, _ = conv2d_fused_ref(x, w, np.ones(64), np.zeros(64), stride=2, pad_lo=0, pad_hi=1)
    ref = conv2d_naive(x.astype(np.float64), w.astype(np.float16).astype(np.float64), 2, 1, 0, 1)
    assert np.abs(y - ref).max() < 1e-9
