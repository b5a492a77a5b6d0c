"""Post-path transform to_orig_cam: oracle and joint tables against vectors produced by the
reference's own volumetric.to_orig_cam / JointInfo (tests/golden/post.npz, joints.npz; oracle/gen_golden.py),
and the CUDA entry point metro_to_orig_cam against both."""
import os

import numpy as np
import pytest

from metro_pose3d_b200.joints import exported_joint_info, model_joint_info
from oracle.metro_oracle import to_orig_cam_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _inputs(ds):
    """Same seeded inputs as oracle/gen_golden.py (synth_rotations and the pose draw)."""
    g = np.load(os.path.join(GOLD, 'post.npz'))
    n, j, pseed, rseed = (int(v) for v in g[f'{ds}_meta'])
    poses = np.random.RandomState(pseed).randn(n, j, 3) * 400.0
    rng = np.random.RandomState(rseed)
    rots = []
    for i in range(n):
        q, r = np.linalg.qr(rng.randn(3, 3))
        q = q * np.sign(np.diag(r))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        if i % 2 == 1:
            q = q @ np.diag([-1.0, 1.0, 1.0])
        rots.append(q)
    return poses, np.stack(rots), g[f'{ds}_orig_cam']


def test_mirror_mapping_matches_the_reference_class():
    g = np.load(os.path.join(GOLD, 'joints.npz'))
    assert model_joint_info('h36m').mirror_mapping == list(g['h36m_model_mirror'])
    assert model_joint_info('merged').mirror_mapping == list(g['merged_model_mirror'])
    assert exported_joint_info('h36m').mirror_mapping == list(g['h36m_export_mirror'])
    mm = model_joint_info('h36m').mirror_mapping
    assert [mm[i] for i in mm] == list(range(len(mm)))          # an involution


@pytest.mark.parametrize('ds', ['h36m', 'merged'])
def test_oracle_matches_reference_code(ds):
    poses, rot, want = _inputs(ds)
    assert (np.linalg.det(rot) > 0).tolist() == [True, False] * (len(rot) // 2)
    got = to_orig_cam_ref(poses, rot, model_joint_info(ds).mirror_mapping)
    assert np.abs(got - want).max() < 1e-9


def test_identity_and_pure_flip_known_answers():
    x = np.arange(17 * 3, dtype=np.float64).reshape(1, 17, 3)
    mm = model_joint_info('h36m').mirror_mapping
    assert np.array_equal(to_orig_cam_ref(x, np.eye(3)[None], mm), x)
    flip = np.diag([-1.0, 1.0, 1.0])[None]
    y = to_orig_cam_ref(x, flip, mm)
    assert np.array_equal(y[0, :, 0], -x[0, mm, 0]) and np.array_equal(y[0, :, 1:], x[0, mm, 1:])


@pytest.mark.gpu
@pytest.mark.parametrize('ds', ['h36m', 'merged'])
def test_cuda_to_orig_cam(ds):
    """float32 on the device against the reference-code vectors: 1e-3 mm on coordinates of ~1e3 mm."""
    import torch
    from metro_pose3d_b200.inference import to_orig_cam
    poses, rot, want = _inputs(ds)
    mm = model_joint_info(ds).mirror_mapping
    got = to_orig_cam(torch.from_numpy(poses.astype(np.float32)).cuda(), torch.from_numpy(rot.astype(np.float32)).cuda(), mm)
    assert np.abs(got.cpu().numpy() - want).max() < 1e-3
    big = torch.randn(4096, len(mm), 3, device='cuda') * 500
    r = torch.from_numpy(rot.astype(np.float32)).cuda().repeat(4096 // len(rot) + 1, 1, 1)[:4096]
    ref = to_orig_cam_ref(big.cpu().numpy(), r.cpu().numpy(), mm)
    assert np.abs(to_orig_cam(big, r, mm).cpu().numpy() - ref).max() < 2e-3


@pytest.mark.gpu
def test_cuda_to_orig_cam_argument_errors():
    import torch
    from metro_pose3d_b200.inference import to_orig_cam
    x = torch.zeros(2, 17, 3, device='cuda')
    r = torch.eye(3, device='cuda').repeat(2, 1, 1)
    with pytest.raises(ValueError):
        to_orig_cam(x, r, list(range(16)))
    with pytest.raises(ValueError):
        to_orig_cam(x, r, [17] * 17)
    with pytest.raises(ValueError):
        to_orig_cam(x, r[:1], list(range(17)))


def test_round_trip_property():
    """Size-independent property: rotating into the original camera and back with the transposed matrix returns
    the skeleton, flipped crops included (the left/right swap is an involution and commutes with the rotation)."""
    rng = np.random.RandomState(3)
    mm = model_joint_info('merged').mirror_mapping
    x = rng.randn(64, len(mm), 3) * 300.0
    rots = []
    for i in range(64):
        q, r = np.linalg.qr(rng.randn(3, 3))
        q = q * np.sign(np.diag(r))
        if (np.linalg.det(q) < 0) != (i % 3 == 0):        # every third matrix is a reflection
            q[:, 0] = -q[:, 0]
        rots.append(q)
    rot = np.stack(rots)
    y = to_orig_cam_ref(x, rot, mm)
    back = to_orig_cam_ref(y, np.transpose(rot, (0, 2, 1)), mm)
    assert np.abs(back - x).max() < 1e-9
