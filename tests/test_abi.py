"""The C-ABI shared library loads without a GPU, exports every symbol include/metro.h declares, and
reports errors the way the reference does (ValueError for bad strides; loud failure with no device)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from metro_pose3d_b200 import lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'metro.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(metro_[a-z0-9_]+)\s*\(', src)))


def test_every_declared_symbol_is_exported(libmetro):
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(libmetro, n), f'{n} declared in include/metro.h but not exported by libmetro.so'
    assert sorted(lib.EXPORTS) == names


def test_version_and_blob_size(libmetro):
    assert b'sm_100a' in libmetro.metro_version()
    spec = lib.make_spec('resnet_v2_101', 16, 19, list(range(19)))
    assert lib.blob_floats(spec) > 42_000_000


def test_value_errors_cross_the_abi_as_status_codes(libmetro):
    bad = lib.make_spec('resnet_v2_50', 10, 17, [0])
    n = C.c_uint64()
    assert libmetro.metro_blob_floats(C.byref(bad), C.byref(n)) == lib.METRO_ERR_VALUE
    assert b'multiple of 4' in libmetro.metro_last_error()
    d = lib.SoftargmaxDesc(8, 17, 8, 32, 1, 256, 2200.0, 1, C.cast((C.c_int32 * 1)(99), C.POINTER(C.c_int32)), 0, 0, 0)
    b = C.c_uint64()
    assert libmetro.metro_softargmax_workspace_bytes(C.byref(d), 4, C.byref(b)) == lib.METRO_ERR_VALUE
    assert b'permutation' in libmetro.metro_last_error()


def test_softargmax_workspace_size(libmetro):
    perm = (C.c_int32 * 17)(*range(17))
    d = lib.SoftargmaxDesc(16, 17, 8, 16, 1, 256, 2200.0, 17, C.cast(perm, C.POINTER(C.c_int32)), 0, 0, 0)
    b = C.c_uint64()
    lib.check(libmetro.metro_softargmax_workspace_bytes(C.byref(d), 256, C.byref(b)))
    assert 1024 <= b.value < 64 << 20


def test_no_gpu_means_loud_failure_not_fallback(libmetro):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from metro_pose3d_b200.inference import MetroModel
    with pytest.raises(lib.MetroError, match='no CPU fallback'):
        MetroModel('resnet_v2_50', 32, 'h36m', max_batch=1)


def test_wrong_blob_size_is_a_value_error(libmetro):
    spec = lib.make_spec('resnet_v2_50', 32, 17, list(range(17)))
    blob = np.zeros(10, np.float32)
    h = C.c_void_p()
    st = libmetro.metro_create(C.byref(spec), blob.ctypes.data_as(C.c_void_p), blob.size, 0, C.byref(h))
    assert st == lib.METRO_ERR_VALUE and b'weight blob' in libmetro.metro_last_error()


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, 'metro_pose3d_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cpp', '.h', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b|#include\s+".*oracle', text, flags=re.M), \
                    f'{f} pulls in the oracle'


def test_argument_errors_of_the_side_entry_points_need_no_gpu(libmetro):
    """The validation of the round-2 entry points happens before any CUDA call: bad arguments come back as
    METRO_ERR_VALUE (the reference's ValueError) on a machine without a GPU too."""
    vp = C.c_void_p
    one = vp(16)                                          # a non-null pointer that is never dereferenced on these paths
    assert libmetro.metro_back_project(one, one, one, -1, 17, 16, 1, 256, C.c_float(2200.0), one, None) == lib.METRO_ERR_VALUE
    assert libmetro.metro_back_project(one, one, one, 2, 0, 16, 1, 256, C.c_float(2200.0), one, None) == lib.METRO_ERR_VALUE
    assert libmetro.metro_back_project(None, one, one, 2, 17, 16, 1, 256, C.c_float(2200.0), one, None) == lib.METRO_ERR_VALUE
    assert libmetro.metro_back_project(one, one, one, 0, 17, 16, 1, 256, C.c_float(2200.0), one, None) == lib.METRO_OK   # empty batch
    src = (lib.CropSrc * 1)()
    src[0].frame_dev, src[0].height, src[0].width, src[0].row_stride_bytes = 16, 480, 640, 100       # stride < 3 * width
    assert libmetro.metro_extract_crops(src, 1, 256, 0, one, None) == lib.METRO_ERR_VALUE
    assert b'source 0' in libmetro.metro_last_error()
    src[0].row_stride_bytes = 1920
    assert libmetro.metro_extract_crops(src, 1, 256, 300, one, None) == lib.METRO_ERR_VALUE             # border not a byte
    assert libmetro.metro_extract_crops(src, 1, 0, 0, one, None) == lib.METRO_ERR_VALUE
    src[0].height = 40000
    assert libmetro.metro_extract_crops(src, 1, 256, 0, one, None) == lib.METRO_ERR_VALUE
    assert b'32767' in libmetro.metro_last_error()
    assert libmetro.metro_extract_crops(src, 0, 256, 0, one, None) == lib.METRO_OK
    perm = (C.c_int32 * 17)(*range(17))
    d = lib.SoftargmaxDesc(0, 17, 8, 16, 1, 256, 2200.0, 17, C.cast(perm, C.POINTER(C.c_int32)), 0, 0, 0, 0)
    assert libmetro.metro_heatmap_z(C.byref(d), one, 2, one, None) == lib.METRO_ERR_VALUE
    d.side = 16
    assert libmetro.metro_heatmap_z(C.byref(d), None, 2, one, None) == lib.METRO_ERR_VALUE
    assert libmetro.metro_softargmax_coords(C.byref(d), one, 2, None, None, one, None) == lib.METRO_ERR_VALUE     # no output at all
    assert libmetro.metro_get_joint_info(None, None, 0, None, None, 0, None, None) == lib.METRO_ERR_VALUE
    assert libmetro.metro_graph_stats(None, None, None) == lib.METRO_ERR_VALUE


def test_spec_validation_of_precision_and_joint_tables(libmetro):
    """metro_create checks the new spec fields before it looks for a device."""
    blob = np.zeros(4, np.float32)
    h = C.c_void_p()
    spec = lib.make_spec('resnet_v2_50', 32, 17, list(range(17)), precision=7)
    assert libmetro.metro_create(C.byref(spec), blob.ctypes.data_as(C.c_void_p), blob.size, 0, C.byref(h)) == lib.METRO_ERR_VALUE
    assert b'precision' in libmetro.metro_last_error()
    spec = lib.make_spec('resnet_v2_50', 32, 17, list(range(17)), joint_names=['a', 'b'])            # 2 names for 17 joints
    assert libmetro.metro_create(C.byref(spec), blob.ctypes.data_as(C.c_void_p), blob.size, 0, C.byref(h)) == lib.METRO_ERR_VALUE
    assert b'joint_names' in libmetro.metro_last_error()
    spec = lib.make_spec('resnet_v2_50', 32, 17, list(range(17)), joint_edges=[(0, 99)])
    assert libmetro.metro_create(C.byref(spec), blob.ctypes.data_as(C.c_void_p), blob.size, 0, C.byref(h)) == lib.METRO_ERR_VALUE
    assert b'joint_edges' in libmetro.metro_last_error()
