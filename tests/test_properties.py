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
