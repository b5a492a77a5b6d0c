"""Property tests (hypothesis) of the decode oracle -- the properties SURVEY 8c asks the multi-GPU path to rely on:
batch-permutation equivariance, N-sharding invariance (a crop's result does not depend on which shard or batch
position it lands in), invariance to a per-(crop, joint) constant, and the root joint decoding to zero."""
import numpy as np
from hypothesis import given, settings, strategies as st

from metro_pose3d_b200.dist import shard_bounds
from metro_pose3d_b200.joints import export_permutation
from oracle.metro_oracle import decode_ref

_PERM = export_permutation('h36m')


def _head(n, side, seed):
    rng = np.random.RandomState(seed)
    return (rng.randn(n, side, side, 8 * 17) * 3.0).astype(np.float32)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 6), side=st.sampled_from([2, 4, 8]), seed=st.integers(0, 10_000), data=st.data())
def test_batch_permutation_equivariance(n, side, seed, data):
    x = _head(n, side, seed)
    order = data.draw(st.permutations(list(range(n))))
    a = decode_ref(x, 17, 256 // side, _PERM)
    b = decode_ref(x[order], 17, 256 // side, _PERM)
    assert np.array_equal(a[order], b)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 9), world=st.integers(1, 4), side=st.sampled_from([4, 8]), seed=st.integers(0, 10_000))
def test_sharding_invariance(n, world, side, seed):
    """Decoding the shards rank by rank (dist.shard_bounds, the split bench.py and dist.py use) and concatenating
    equals decoding the whole batch, bit for bit."""
    x = _head(n, side, seed)
    whole = decode_ref(x, 17, 256 // side, _PERM)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        if hi > lo:
            parts.append(decode_ref(x[lo:hi], 17, 256 // side, _PERM))
    assert np.array_equal(np.concatenate(parts), whole)


@settings(max_examples=25, deadline=None)
@given(side=st.sampled_from([4, 8]), seed=st.integers(0, 10_000), shift=st.floats(-50, 50))
def test_constant_shift_per_joint_and_root_is_zero(side, seed, shift):
    x = _head(2, side, seed)
    a = decode_ref(x, 17, 256 // side, _PERM)
    y = x.astype(np.float64).reshape(2, side, side, 8, 17)
    y[..., 3] += shift                                   # every voxel of joint 3: the softmax does not see it
    b = decode_ref(y.reshape(2, side, side, 136), 17, 256 // side, _PERM)
    assert np.abs(a - b).max() < 1e-6
    assert np.all(a[:, _PERM.index(16)] == 0)            # the pelvis (last model joint) is the origin


# ---- crop extraction oracle (oracle/crop_oracle.py: reproject_image_fast = homography + cv2.remap arithmetic) ----
from oracle.crop_oracle import reproject_image_fast_ref, remap_bilinear_u8


@settings(max_examples=25, deadline=None)
@given(h=st.integers(8, 40), w=st.integers(8, 40), dx=st.integers(-6, 6), dy=st.integers(-6, 6), seed=st.integers(0, 10_000))
def test_integer_translation_is_an_exact_shift_with_constant_border(h, w, dx, dy, seed):
    """A homography that translates by whole pixels copies pixels (the single non-zero weight is 32767 / 32768 of OpenCV's
    int16 table, which still rounds to the pixel) and fills what falls outside the frame with the border value."""
    frame = np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)
    hm = np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]], np.float32)
    got = reproject_image_fast_ref(frame, hm, h, w, border_value=9)
    yy, xx = np.mgrid[:h, :w]
    sy, sx = yy + dy, xx + dx
    inside = (sy >= 0) & (sy < h) & (sx >= 0) & (sx < w)
    want = np.where(inside[..., None], frame[np.clip(sy, 0, h - 1), np.clip(sx, 0, w - 1)], 9).astype(np.uint8)
    assert np.array_equal(got, want)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), scale=st.floats(0.3, 3.0))
def test_constant_image_stays_constant_inside_the_frame(seed, scale):
    """Bilinear weights sum to 2^15 (or 2^15 - 1 on an exact pixel): a constant frame resamples to the same constant wherever
    all four taps are inside, whatever the (here: scaling) homography."""
    rng = np.random.RandomState(seed)
    c = int(rng.randint(0, 256))
    frame = np.full((32, 32, 3), c, np.uint8)
    hm = np.array([[scale, 0, 3.3], [0, scale, 2.1], [0, 0, 1]], np.float32)
    mx = (np.arange(8, dtype=np.float32) * np.float32(scale) + np.float32(3.3))[None, :].repeat(8, 0)
    my = (np.arange(8, dtype=np.float32) * np.float32(scale) + np.float32(2.1))[:, None].repeat(8, 1)
    got = reproject_image_fast_ref(frame, hm, 8, 8)
    inside = (mx >= 0) & (mx <= 30.9) & (my >= 0) & (my <= 30.9)
    assert np.all(got[inside] == c)
    assert np.array_equal(got, remap_bilinear_u8(frame, mx.astype(np.float32), my.astype(np.float32)))


# ---- absolute-scale variant (oracle true_root_depth_ref: volumetric.py:190-198,285) ----
from oracle.metro_oracle import true_root_depth_ref


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), stride=st.sampled_from([4, 8, 16, 32]), j=st.sampled_from([17, 19, 53]))
def test_back_projection_round_trip(seed, stride, j):
    """Projecting the back-projected joints with the camera's intrinsics gives back the predicted image coordinates
    (heatmap_to_image of the 2D part), the root joint sits at the given depth, and depths relative to the root are the
    heatmap's z differences in millimetres -- for any camera, stride and joint count."""
    rng = np.random.RandomState(seed)
    n = 3
    c = rng.rand(n, j, 3)
    f = rng.uniform(200.0, 1500.0, n)
    k = np.zeros((n, 3, 3))
    k[:, 0, 0], k[:, 1, 1], k[:, 2, 2] = f, f * rng.uniform(0.9, 1.1, n), 1.0
    k[:, 0, 2], k[:, 1, 2] = rng.uniform(100, 156, n), rng.uniform(100, 156, n)
    z = rng.uniform(1500.0, 8000.0, n)
    x = true_root_depth_ref(c, np.linalg.inv(k), z, stride)
    proj = np.einsum('bij,bcj->bci', k, x / x[..., 2:3])
    lrc = 255 - (255 % stride) - 1
    assert np.abs(proj[..., :2] - (c[..., :2] * lrc + stride // 2)).max() < 1e-8
    assert np.abs(x[:, -1, 2] - z).max() < 1e-9
    assert np.abs((x[..., 2] - x[:, -1:, 2]) - (c[..., 2] - c[:, -1:, 2]) * 2200.0).max() < 1e-9
