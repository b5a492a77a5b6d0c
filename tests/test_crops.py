"""Crop extraction: cameralib.reproject_image_fast (src/cameralib.py:406-429) = homography + cv2.remap.

CPU: the oracle's restatement of cv2.remap's fixed-point bilinear arithmetic against cv2 itself (bit for bit), and against
the vectors the reference's own function produced (tests/golden/crops.npz, oracle/gen_golden.py).  GPU: metro_extract_crops
through the C-ABI against both, bit for bit (except the handful of pixels where the reference's BLAS rounds a float32
coordinate differently from the fixed evaluation order, recorded in the fixture)."""
import os

import numpy as np
import pytest

from oracle.crop_oracle import crop_homography, remap_bilinear_u8, reproject_image_fast_ref, source_coords
from oracle.gen_golden import synth_crop_cameras, synth_frame

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = ['vga', 'tall']


def test_remap_restatement_equals_cv2_bit_for_bit():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(0)
    for trial in range(12):
        h, w = int(rng.integers(40, 300)), int(rng.integers(40, 400))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        mx = (rng.random((128, 128), dtype=np.float32) * (w + 20) - 10).astype(np.float32)
        my = (rng.random((128, 128), dtype=np.float32) * (h + 20) - 10).astype(np.float32)
        if trial % 3 == 0:                                 # exact integer / quarter-pixel coordinates, far outside too
            mx, my = np.round(mx * 4) / 4, np.round(my * 8) / 8
            mx[0, :4] = [-1e9, 1e9, np.inf, -5.0]
        for border in (0, 77):
            ref = cv2.remap(img, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=(border,) * 3)
            assert np.array_equal(remap_bilinear_u8(img, mx, my, border), ref), (trial, border)


@pytest.mark.parametrize('name', CASES)
def test_oracle_against_the_reference_function(name):
    g = np.load(os.path.join(GOLD, 'crops.npz'))
    h, w, side, n, seed = (int(v) for v in g[f'{name}_meta'])
    frame = synth_frame(h, w, seed)
    cams = synth_crop_cameras(n, h, w, seed + 1, side)
    hs = np.stack([crop_homography(*c) for c in cams])
    assert np.array_equal(hs, g[f'{name}_homographies'])
    mine = np.stack([reproject_image_fast_ref(frame, hm, side, side) for hm in hs])
    diff = np.argwhere(np.any(mine != g[f'{name}_crops'], axis=-1))
    assert np.array_equal(diff, g[f'{name}_blas_mismatch']) and len(diff) <= 1e-3 * mine[..., 0].size
    # those pixels differ because a float32 source coordinate lands in the neighbouring 1/32-pixel cell: by a few grey levels
    assert np.abs(mine.astype(int) - g[f'{name}_crops'].astype(int)).max() <= 16
    assert 0.05 < (g[f'{name}_crops'].sum(-1) == 0).mean() < 0.6            # the constant border takes part


def test_identity_homography_and_half_pixel_shift():
    frame = synth_frame(64, 64, 3)
    eye = np.eye(3, dtype=np.float32)
    assert np.array_equal(reproject_image_fast_ref(frame, eye, 64, 64), frame)
    shift = eye.copy(); shift[0, 2] = 0.5                                     # samples halfway between columns x and x + 1
    got = reproject_image_fast_ref(frame, shift, 64, 63)
    want = (frame[:, :-1].astype(int) + frame[:, 1:].astype(int) + 1) >> 1   # weights 16384 / 16384, + 2^14 >> 15
    assert np.array_equal(got, want.astype(np.uint8))
    mx, my = source_coords(shift, 4, 4)
    assert mx.dtype == np.float32 and np.array_equal(mx[0], [0.5, 1.5, 2.5, 3.5]) and np.array_equal(my[:, 0], [0, 1, 2, 3])


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_cuda_crops_bit_exact(name):
    import torch
    from metro_pose3d_b200.inference import extract_crops
    g = np.load(os.path.join(GOLD, 'crops.npz'))
    h, w, side, n, seed = (int(v) for v in g[f'{name}_meta'])
    frame = synth_frame(h, w, seed)
    hs = g[f'{name}_homographies']
    got = extract_crops(torch.from_numpy(frame).cuda(), hs, side=side).cpu().numpy()
    mine = np.stack([reproject_image_fast_ref(frame, hm, side, side) for hm in hs])
    assert np.array_equal(got, mine)                                          # the oracle: every pixel, bit for bit
    diff = np.argwhere(np.any(got != g[f'{name}_crops'], axis=-1))           # the reference's own output
    assert np.array_equal(diff, g[f'{name}_blas_mismatch'])
    got77 = extract_crops(torch.from_numpy(frame).cuda(), hs, side=side, border_value=77).cpu().numpy()
    assert np.array_equal(got77, np.stack([reproject_image_fast_ref(frame, hm, side, side, 77) for hm in hs]))


@pytest.mark.gpu
def test_cuda_crops_feed_the_network():
    """Frames -> crops -> poses without leaving the device; one frame per crop; 256 x 256 crops; more crops than one launch
    carries (32); argument errors."""
    import torch
    from metro_pose3d_b200.inference import MetroModel, extract_crops
    n = 40
    frames = [torch.from_numpy(synth_frame(300 + 7 * i, 420 - 3 * i, 100 + i)).cuda() for i in range(n)]
    cams = [synth_crop_cameras(1, 300 + 7 * i, 420 - 3 * i, 200 + i, 256)[0] for i in range(n)]
    hs = np.stack([crop_homography(*c) for c in cams])
    crops = extract_crops(frames, hs, side=256)
    assert crops.shape == (n, 256, 256, 3) and crops.dtype == torch.uint8
    for i in (0, 31, 32, 39):
        want = reproject_image_fast_ref(frames[i].cpu().numpy(), hs[i], 256, 256)
        assert np.array_equal(crops[i].cpu().numpy(), want), i
    model = MetroModel('resnet_v2_50', 32, 'h36m', max_batch=n)
    poses = model.infer(crops)
    assert poses.shape == (n, 17, 3) and torch.isfinite(poses).all()
    assert torch.equal(poses, model.infer(crops.float().div(255.0)))         # uint8 ingestion == improc.normalize01 feed
    with pytest.raises(ValueError):
        extract_crops(frames[:3], hs[:4])
    with pytest.raises(ValueError):
        extract_crops(frames[0].float(), hs[:1])
