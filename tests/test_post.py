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


# ---- absolute-scale variant and the depth marginal (second half) ------------------------------

def _trd_inputs(ds):
    import oracle.gen_golden as G                     # the seeded input generators only (no reference needed)
    j = {'h36m': 17, 'merged': 53}[ds]
    rng = np.random.RandomState(23 + j)
    coords01 = rng.rand(6, j, 3)
    return coords01, G.synth_inv_intrinsics(6, 29 + j), rng.uniform(2000.0, 6000.0, 6)


@pytest.mark.parametrize('ds', ['h36m', 'merged'])
def test_true_root_depth_oracle_matches_reference_code(ds):
    from oracle.metro_oracle import true_root_depth_ref
    g = np.load(os.path.join(GOLD, 'post.npz'))
    c, k, z = _trd_inputs(ds)
    for stride in (4, 16, 32):
        a = true_root_depth_ref(c, k, z, stride)
        assert np.abs(a - g[f'{ds}_trd_abs_s{stride}']).max() < 1e-9
        assert np.abs((a - a[:, -1:]) - g[f'{ds}_trd_rel_s{stride}']).max() < 1e-9


def test_heatmap_pred_z_oracle_matches_reference_code():
    from oracle.metro_oracle import heatmap_pred_z_ref
    from metro_pose3d_b200.weights import synth_head
    g = np.load(os.path.join(GOLD, 'post.npz'))
    for name in ('A', 'C'):
        n, side, j, seed = (int(v) for v in g[f'pred_z_{name}_meta'])
        got = heatmap_pred_z_ref(synth_head(n, side, j, seed=seed), j)
        assert np.abs(got - g[f'pred_z_{name}']).max() < 1e-12
        assert np.allclose(got.sum(-1), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize('ds', ['h36m', 'merged'])
def test_cuda_back_project(ds):
    """metro_back_project against the reference-code vectors: 1e-3 mm on identical (float32) inputs, 2e-3 mm against the
    float64-input vectors (the float32 rounding of the inputs alone moves a 6 m ray by up to ~5e-4 mm)."""
    import torch
    from metro_pose3d_b200.inference import back_project
    from oracle.metro_oracle import true_root_depth_ref
    g = np.load(os.path.join(GOLD, 'post.npz'))
    c, k, z = (a.astype(np.float32) for a in _trd_inputs(ds))
    for stride in (4, 16, 32):
        got = back_project(torch.from_numpy(c).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(z).cuda(), stride).cpu().numpy()
        assert np.abs(got - true_root_depth_ref(c, k, z, stride)).max() < 1e-3
        assert np.abs(got - g[f'{ds}_trd_abs_s{stride}']).max() < 2e-3


@pytest.mark.gpu
def test_cuda_coords_and_depth_marginal():
    """The decode's second fetch (heatmap coordinates in [0,1]) against the reference-code vectors of tests/golden/decode.npz
    and the depth marginal t.heatmap_pred_z against post.npz."""
    import torch
    from metro_pose3d_b200.inference import SoftArgmax
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.weights import synth_head
    d = np.load(os.path.join(GOLD, 'decode.npz'))
    for case, ds in (('A', 'h36m'), ('B', 'h36m'), ('C', 'coco19'), ('E', 'coco19')):
        n, side, stride, j, seed = (int(v) for v in d[f'{case}_meta'])
        op = SoftArgmax(side, j, stride, export_permutation(ds))
        poses, c01 = op.coords(torch.from_numpy(synth_head(n, side, j, seed=seed)).cuda())
        assert np.abs(c01.cpu().numpy() - d[f'{case}_coords01']).max() < 5e-7          # 1e-3 mm / 2200 mm
        assert np.abs(poses.cpu().numpy() - d[f'{case}_poses']).max() < 1e-3
    g = np.load(os.path.join(GOLD, 'post.npz'))
    for name in ('A', 'C'):
        n, side, j, seed = (int(v) for v in g[f'pred_z_{name}_meta'])
        op = SoftArgmax(side, j, 256 // side, list(range(j)))
        got = op.heatmap_z(torch.from_numpy(synth_head(n, side, j, seed=seed)).cuda()).cpu().numpy()
        assert np.abs(got - g[f'pred_z_{name}']).max() < 1e-6


@pytest.mark.gpu
def test_cuda_network_coords_feed_back_projection():
    """images -> metro_infer_coords -> metro_back_project: the 'true-root-depth' evaluation path end to end, strict
    precision against the float64 oracle (1e-3 mm), and the tensor-core path's coords consistent with its poses."""
    import torch
    from metro_pose3d_b200.inference import MetroModel, back_project
    from metro_pose3d_b200.joints import export_permutation
    from metro_pose3d_b200.spec import NetSpec
    from metro_pose3d_b200.weights import synth_images, synth_weights
    from oracle.metro_oracle import OracleNet, decode_ref, true_root_depth_ref
    import oracle.gen_golden as G
    spec = NetSpec('resnet_v2_50', 32, 17)
    w = synth_weights(spec, 0)
    perm = export_permutation('h36m')
    img = synth_images(2, seed=31)
    x = torch.from_numpy(img).cuda()
    head = OracleNet(spec, w, perm, 'fp64').forward_head(img)
    want_c = decode_ref(head, 17, 32, perm, return_coords01=True)
    inv_k = G.synth_inv_intrinsics(2, 5).astype(np.float32)
    z = np.array([3000.0, 4500.0], np.float32)
    strict = MetroModel('resnet_v2_50', 32, 'h36m', weights=w, max_batch=2, precision='strict')
    poses, c01 = strict.infer_coords(x)
    assert np.abs(c01.cpu().numpy() - want_c).max() < 5e-7
    absolute = back_project(c01, torch.from_numpy(inv_k).cuda(), torch.from_numpy(z).cuda(), 32).cpu().numpy()
    assert np.abs(absolute - true_root_depth_ref(want_c, inv_k, z, 32)).max() < 2e-3
    fast = MetroModel('resnet_v2_50', 32, 'h36m', weights=w, max_batch=2)
    p2, c2 = fast.infer_coords(x)
    assert np.array_equal(p2.cpu().numpy(), fast.infer(x).cpu().numpy())
    # the coords are the poses before metric scaling / root subtraction / permutation (volumetric.py:303-306)
    c2 = c2.cpu().numpy().astype(np.float64)
    lrc = 255 - (255 % 32) - 1
    metric = np.concatenate([c2[..., :2] * lrc * 2200.0 / 256, c2[..., 2:] * 2200.0], -1)
    assert np.abs((metric - metric[:, -1:])[:, perm] - p2.cpu().numpy()).max() < 2e-3
