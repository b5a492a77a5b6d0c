"""ctypes binding of libmetro.so (include/metro.h).  No compute happens in Python.

The library is built in-tree by ``metro_pose3d_b200/csrc/build.sh`` (``__graft_entry__.build()``).
If it is missing the import of the product path fails loudly: there is no CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmetro.so')

METRO_OK, METRO_ERR_VALUE, METRO_ERR_CUDA, METRO_ERR_NO_DEVICE, METRO_ERR_NOMEM, METRO_ERR_INTERNAL = range(6)
METRO_F32, METRO_F16 = 0, 1
METRO_PREC_F16, METRO_PREC_STRICT, METRO_PREC_STRICT_F16 = 0, 1, 2
PRECISIONS = {'f16': METRO_PREC_F16, 'strict': METRO_PREC_STRICT, 'strict_f16': METRO_PREC_STRICT_F16}

EXPORTS = [
    'metro_last_error', 'metro_version', 'metro_blob_floats', 'metro_plan_describe', 'metro_create',
    'metro_destroy', 'metro_workspace_bytes', 'metro_get_joint_info', 'metro_infer', 'metro_infer_host', 'metro_infer_u8', 'metro_infer_host_u8', 'metro_to_orig_cam',
    'metro_softargmax_workspace_bytes', 'metro_softargmax', 'metro_softargmax_coords', 'metro_infer_coords', 'metro_heatmap_z',
    'metro_back_project', 'metro_extract_crops', 'metro_conv2d', 'metro_debug_read',
    'metro_profile', 'metro_launch_count', 'metro_graph_stats',
]


class MetroSpec(C.Structure):
    _fields_ = [
        ('arch', C.c_int32), ('stride', C.c_int32), ('n_joints_model', C.c_int32), ('depth', C.c_int32),
        ('centered_stride', C.c_int32), ('proc_side', C.c_int32), ('box_size_mm', C.c_float),
        ('n_joints_out', C.c_int32), ('permutation', C.POINTER(C.c_int32)), ('max_batch', C.c_int32),
        ('head_dtype', C.c_int32), ('keep_activations', C.c_int32), ('precision', C.c_int32),
        ('joint_names', C.c_char_p), ('n_joint_edges', C.c_int32), ('joint_edges', C.POINTER(C.c_int32)),
    ]


class SoftargmaxDesc(C.Structure):
    _fields_ = [
        ('side', C.c_int32), ('n_joints_model', C.c_int32), ('depth', C.c_int32), ('stride', C.c_int32),
        ('centered_stride', C.c_int32), ('proc_side', C.c_int32), ('box_size_mm', C.c_float),
        ('n_joints_out', C.c_int32), ('permutation', C.POINTER(C.c_int32)), ('head_dtype', C.c_int32),
        ('splits', C.c_int32), ('lanes', C.c_int32), ('word_bytes', C.c_int32),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ('n', C.c_int32), ('in_side', C.c_int32), ('cin', C.c_int32), ('cout', C.c_int32), ('k', C.c_int32),
        ('stride', C.c_int32), ('rate', C.c_int32), ('pad_lo', C.c_int32), ('relu', C.c_int32),
        ('out_dtype', C.c_int32), ('res_stride', C.c_int32), ('res_shift', C.c_int32), ('cin2', C.c_int32),
    ]


class CropSrc(C.Structure):
    _fields_ = [('frame_dev', C.c_void_p), ('height', C.c_int32), ('width', C.c_int32), ('row_stride_bytes', C.c_int32),
                ('homography', C.c_float * 9)]


class MetroError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def build_library(verbose: bool = False) -> str:
    """Compile libmetro.so for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(['bash', os.path.join(_HERE, 'csrc', 'build.sh')], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise MetroError('building libmetro.so failed')
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MetroError(
            f'{LIB_PATH} not found: run `python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU fallback for the product path)')
    lib = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    lib.metro_last_error.restype = C.c_char_p
    lib.metro_version.restype = C.c_char_p
    lib.metro_blob_floats.argtypes = [C.POINTER(MetroSpec), C.POINTER(u64)]
    lib.metro_plan_describe.argtypes = [C.POINTER(MetroSpec), C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.metro_create.argtypes = [C.POINTER(MetroSpec), vp, u64, i32, C.POINTER(vp)]
    lib.metro_destroy.argtypes = [vp]
    lib.metro_workspace_bytes.argtypes = [vp, i32, C.POINTER(u64)]
    lib.metro_get_joint_info.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(i32), i32,
                                         C.POINTER(i32), C.POINTER(i32)]
    lib.metro_infer.argtypes = [vp, vp, i32, vp, vp]
    lib.metro_infer_u8.argtypes = [vp, vp, i32, vp, vp]
    lib.metro_infer_host.argtypes = [vp, vp, i32, vp]
    lib.metro_infer_host_u8.argtypes = [vp, vp, i32, vp]
    lib.metro_to_orig_cam.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.metro_softargmax_workspace_bytes.argtypes = [C.POINTER(SoftargmaxDesc), i32, C.POINTER(u64)]
    lib.metro_softargmax.argtypes = [C.POINTER(SoftargmaxDesc), vp, i32, vp, vp, vp]
    lib.metro_softargmax_coords.argtypes = [C.POINTER(SoftargmaxDesc), vp, i32, vp, vp, vp, vp]
    lib.metro_infer_coords.argtypes = [vp, vp, i32, vp, vp, vp]
    lib.metro_heatmap_z.argtypes = [C.POINTER(SoftargmaxDesc), vp, i32, vp, vp]
    lib.metro_back_project.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, C.c_float, vp, vp]
    lib.metro_extract_crops.argtypes = [C.POINTER(CropSrc), i32, i32, i32, vp, vp]
    lib.metro_conv2d.argtypes = [C.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    lib.metro_debug_read.argtypes = [vp, C.c_char_p, vp, u64, C.POINTER(u64)]
    lib.metro_profile.argtypes = [vp, vp, i32, vp, vp, C.c_char_p, C.c_size_t, C.POINTER(i32)]
    lib.metro_launch_count.argtypes = [vp, i32, C.POINTER(i32)]
    lib.metro_graph_stats.argtypes = [vp, C.POINTER(i32), C.POINTER(C.c_int64)]
    for name in EXPORTS:
        if name not in ('metro_last_error', 'metro_version'):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(status: int):
    """Map status codes to the exceptions the reference raises for the same conditions."""
    if status == METRO_OK:
        return
    msg = load().metro_last_error().decode('utf-8', 'replace')
    if status == METRO_ERR_VALUE:
        raise ValueError(msg)
    if status == METRO_ERR_NOMEM:
        raise MemoryError(msg)
    raise MetroError(f'[{status}] {msg}')


def make_spec(arch: str, stride: int, n_joints_model: int, permutation: Sequence[int], max_batch: int = 1,
              depth: int = 8, centered_stride: bool = True, proc_side: int = 256, box_size_mm: float = 2200.0,
              head_dtype: int = METRO_F32, keep_activations: bool = False, precision: int = METRO_PREC_F16,
              joint_names: Optional[Sequence[str]] = None, joint_edges=None):
    archs = {'resnet_v2_50': 50, 'resnet_v2_101': 101}
    if arch not in archs:
        raise ValueError(f'unknown architecture {arch!r}')
    perm = (C.c_int32 * len(permutation))(*permutation)
    names = None if joint_names is None else '\n'.join(joint_names).encode()
    flat = [] if joint_edges is None else [int(v) for e in joint_edges for v in e]
    edges = (C.c_int32 * max(len(flat), 1))(*flat)
    spec = MetroSpec(archs[arch], stride, n_joints_model, depth, int(centered_stride), proc_side,
                     box_size_mm, len(permutation), C.cast(perm, C.POINTER(C.c_int32)), max_batch, head_dtype,
                     int(keep_activations), precision, names, len(flat) // 2,
                     C.cast(edges, C.POINTER(C.c_int32)) if flat else None)
    spec._keepalive = (perm, names, edges)
    return spec


def plan_describe(spec: MetroSpec) -> dict:
    lib = load()
    need = C.c_size_t(0)
    check(lib.metro_plan_describe(C.byref(spec), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    check(lib.metro_plan_describe(C.byref(spec), buf, need.value, None))
    return json.loads(buf.value.decode())


def blob_floats(spec: MetroSpec) -> int:
    n = C.c_uint64(0)
    check(load().metro_blob_floats(C.byref(spec), C.byref(n)))
    return n.value
