"""CPU oracle of the crop extraction that precedes the hot path.  TEST INFRASTRUCTURE.

Restates ``cameralib.reproject_image_fast`` (src/cameralib.py:406-429): a homography from the two pinhole cameras, the
source coordinate of every output pixel, and OpenCV's ``cv2.remap(..., INTER_LINEAR, BORDER_CONSTANT)`` for uint8
images -- whose arithmetic is fixed point (opencv imgproc/src/imgwarp.cpp, remapBilinear<FixedPtCast<int, uchar, 15>>):
coordinates are rounded to 1/32 pixel, the four bilinear weights are 15-bit integers.

PIN STATUS: pinned against the third-party routine itself: cv2 (4.13, the version in this image) is importable here, so
tests/test_crops.py compares ``remap_bilinear_u8`` with ``cv2.remap`` bit for bit on random images and maps (CPU test)
and oracle/gen_golden.py commits vectors produced by the reference's own ``reproject_image_fast`` (its cameralib
module executed here) for the GPU box, where /root/reference does not exist.

The one thing the restatement fixes that the reference leaves to its BLAS is the order of the three float32
multiply-adds in ``homography @ coords`` (numpy calls sgemm): here it is ((h0*x + h1*y) + h2) with every operation
rounded to float32, no fused multiply-add -- which the CUDA kernel reproduces exactly.  The vectors in
tests/golden/crops.npz record how many pixels of the reference's own output (whatever its BLAS did) differ.
"""
from __future__ import annotations

import numpy as np

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15


def crop_homography(old_intrinsics, old_r, new_intrinsics, new_r) -> np.ndarray:
    """src/cameralib.py:411-413: maps output (new camera) pixels to input (old camera) pixels; float32 [3,3]."""
    old_matrix = np.asarray(old_intrinsics, np.float64) @ np.asarray(old_r, np.float64)
    new_matrix = np.asarray(new_intrinsics, np.float64) @ np.asarray(new_r, np.float64)
    return np.linalg.solve(new_matrix.T, old_matrix.T).T.astype(np.float32)


def source_coords(homography: np.ndarray, out_h: int, out_w: int):
    """src/cameralib.py:415-418 with the float32 evaluation order fixed (module docstring): map_x, map_y float32."""
    h = np.asarray(homography, np.float32)
    y, x = np.mgrid[:out_h, :out_w].astype(np.float32)
    rows = [((h[i, 0] * x).astype(np.float32) + (h[i, 1] * y).astype(np.float32)).astype(np.float32) + h[i, 2] for i in range(3)]
    rows = [r.astype(np.float32) for r in rows]
    return (rows[0] / rows[2]).astype(np.float32), (rows[1] / rows[2]).astype(np.float32)


def remap_bilinear_u8(image: np.ndarray, map_x: np.ndarray, map_y: np.ndarray, border_value: int = 0) -> np.ndarray:
    """cv2.remap(image, map_x, map_y, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=border_value) for
    uint8 images [H, W, C] and float32 maps, in OpenCV's fixed-point arithmetic."""
    img = np.asarray(image)
    assert img.dtype == np.uint8 and img.ndim == 3
    hh, ww, _ = img.shape
    # coordinates -> 1/32 pixel, round half to even (cvRound / cvtps2dq), integer part saturated to int16
    def fixed(m):
        t = np.asarray(m, np.float32) * np.float32(INTER_TAB_SIZE)
        ok = (t >= -2.0 ** 31) & (t < 2.0 ** 31)         # x86 cvtps2dq: NaN / out of range -> INT_MIN
        with np.errstate(invalid='ignore'):
            return np.where(ok, np.rint(np.where(ok, t, 0)).astype(np.int64), -2 ** 31)
    sx, sy = fixed(map_x), fixed(map_y)
    fx, fy = sx & (INTER_TAB_SIZE - 1), sy & (INTER_TAB_SIZE - 1)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    # 15-bit weights of the 2x2 neighbourhood (initInterTab2D): exact multiples of 32; the entry for an integer
    # coordinate would be 32768 and saturates to 32767 (int16)
    w00 = np.minimum((INTER_TAB_SIZE - fx) * (INTER_TAB_SIZE - fy) * 32, 32767)
    w01 = fx * (INTER_TAB_SIZE - fy) * 32
    w10 = (INTER_TAB_SIZE - fx) * fy * 32
    w11 = fx * fy * 32

    def fetch(yy, xx):
        inside = (yy >= 0) & (yy < hh) & (xx >= 0) & (xx < ww)
        v = img[np.clip(yy, 0, hh - 1), np.clip(xx, 0, ww - 1)].astype(np.int64)
        return np.where(inside[..., None], v, border_value)

    acc = (fetch(iy, ix) * w00[..., None] + fetch(iy, ix + 1) * w01[..., None] +
           fetch(iy + 1, ix) * w10[..., None] + fetch(iy + 1, ix + 1) * w11[..., None])
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def reproject_image_fast_ref(image, homography, out_h: int, out_w: int, border_value: int = 0) -> np.ndarray:
    mx, my = source_coords(homography, out_h, out_w)
    return remap_bilinear_u8(image, mx, my, border_value)
