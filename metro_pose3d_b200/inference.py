"""Host-side mirror of the reference's inference contract on top of libmetro.so.

Reference: ``inference.py:31-43`` --

    poses, edges, joint_names = estimate_pose(images_tensor, model_path)

where the frozen graph maps 'input:0' (float32 NHWC [N,256,256,3] in [0,1]) to 'output'
(float32 [N,J,3] root-relative millimetres) and carries the 'joint_edges' / 'joint_names' constants
(src/main.py:106-161).  Here a ``MetroModel`` plays the role of the frozen graph; torch tensors are
used only as device-buffer containers (``data_ptr()``) and for the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import lib as _lib
from .joints import exported_joint_info, export_permutation, model_joint_info
from .spec import NetSpec
from .weights import blob_size, pack_blob, synth_weights


def _torch():
    import torch
    return torch


class MetroModel:
    """One exported model resident on one GPU (== the imported GraphDef + its weight constants)."""

    def __init__(self, arch: str = 'resnet_v2_50', stride: int = 16, dataset: str = 'h36m',
                 weights=None, max_batch: int = 256, device: int = 0, head_dtype: str = 'f32',
                 keep_activations: bool = False, seed: int = 0, n_joints_model: Optional[int] = None,
                 permutation: Optional[Sequence[int]] = None, precision: str = 'f16', joint_info=None,
                 centered_stride: bool = True):
        """``precision``: 'f16' = the tensor-core path (the reference's default float16 export, src/options.py:73);
        'strict' = float64 on CUDA cores (tighter than the reference's ``--dtype float32`` export; slow, for
        verification); 'strict_f16' = float64 arithmetic with the float16 graph's storage roundings."""
        self.lib = _lib.load()
        self.dataset = dataset
        ji = model_joint_info(dataset)
        self.n_joints_model = n_joints_model or ji.n_joints
        self.permutation = list(permutation) if permutation is not None else export_permutation(dataset)
        if joint_info is None and permutation is None:
            joint_info = exported_joint_info(dataset)
        self.spec = NetSpec(arch, stride, self.n_joints_model, centered_stride=centered_stride)   # raises ValueError like the reference
        self.device = device
        self.max_batch = max_batch
        self.precision = precision
        self.head_dtype = {'f32': _lib.METRO_F32, 'f16': _lib.METRO_F16}[head_dtype]
        self._cspec = _lib.make_spec(arch, stride, self.n_joints_model, self.permutation, max_batch,
                                     head_dtype=self.head_dtype, keep_activations=keep_activations,
                                     precision=_lib.PRECISIONS[precision], centered_stride=centered_stride,
                                     joint_names=None if joint_info is None else list(joint_info.names),
                                     joint_edges=None if joint_info is None else list(joint_info.edges))
        if weights is None:
            weights = synth_weights(self.spec, seed)
        blob = weights if isinstance(weights, np.ndarray) else pack_blob(self.spec, weights)
        blob = np.ascontiguousarray(blob, dtype=np.float32)
        if blob.size != blob_size(self.spec):
            raise ValueError(f'weight blob has {blob.size} floats, the model needs {blob_size(self.spec)}')
        h = C.c_void_p()
        _lib.check(self.lib.metro_create(C.byref(self._cspec), blob.ctypes.data_as(C.c_void_p), blob.size,
                                         device, C.byref(h)))
        self._h = h

    @classmethod
    def from_frozen_graph(cls, path_or_bytes, max_batch: int = 256, device: int = 0, head_dtype: str = 'f32',
                          keep_activations: bool = False, precision: str = 'f16') -> 'MetroModel':
        """Loads a model exported by the reference's ``main.export()`` (src/main.py:106-160): the binary GraphDef
        is read without TensorFlow (``pb_import``), its constants become the weight blob, its ``output`` gather
        indices the permutation and its ``joint_names`` / ``joint_edges`` constants the joint tables."""
        from .joints import JointInfo
        from .pb_import import import_frozen_graph, load_frozen_model
        fm = import_frozen_graph(path_or_bytes) if isinstance(path_or_bytes, (bytes, bytearray)) \
            else load_frozen_model(path_or_bytes)
        ji = JointInfo(list(fm.joint_names), [tuple(int(v) for v in e) for e in fm.joint_edges])
        model = cls(fm.arch, fm.stride, dataset='h36m', weights=fm.weights, max_batch=max_batch, device=device,
                    head_dtype=head_dtype, keep_activations=keep_activations, n_joints_model=fm.n_joints_model,
                    permutation=fm.permutation, precision=precision, joint_info=ji, centered_stride=fm.centered_stride)
        model.dataset = 'frozen-graph'
        return model

    # -- the three fetches of the frozen graph ----------------------------------------------------
    def _joint_info(self):
        """'joint_names' / 'joint_edges' through the C-ABI (metro_get_joint_info): the handle holds the tables it was
        created with, whether they came from a dataset or from the constants of an imported .pb."""
        need, ne, nj = C.c_size_t(0), C.c_int32(0), C.c_int32(0)
        _lib.check(self.lib.metro_get_joint_info(self._h, None, 0, C.byref(need), None, 0, C.byref(ne), C.byref(nj)))
        buf = C.create_string_buffer(need.value)
        edges = (C.c_int32 * max(2 * ne.value, 1))()
        _lib.check(self.lib.metro_get_joint_info(self._h, buf, need.value, None, edges, ne.value, None, None))
        names = buf.value.decode().split('\n') if need.value > 1 else []
        return names, np.asarray(list(edges)[:2 * ne.value], dtype=np.int64).reshape(-1, 2)

    @property
    def joint_names(self):
        return self._joint_info()[0]

    @property
    def joint_edges(self) -> np.ndarray:
        return self._joint_info()[1]

    @property
    def n_joints_out(self) -> int:
        return len(self.permutation)

    def infer(self, images, out=None, stream: Optional[int] = None):
        """images: CUDA float32 (or uint8) tensor NHWC [n,256,256,3]; returns CUDA float32 [n,J,3].
        Asynchronous on the current torch stream."""
        torch = _torch()
        if images.dim() != 4 or tuple(images.shape[1:]) != (self.spec.proc_side, self.spec.proc_side, 3):
            raise ValueError(f'expected [N,{self.spec.proc_side},{self.spec.proc_side},3] NHWC, got {tuple(images.shape)}')
        if images.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f'expected float32 or uint8 images, got {images.dtype}')
        if not images.is_cuda or images.device.index != self.device:
            raise ValueError(f'images must live on cuda:{self.device}')
        images = images.contiguous()
        n = images.shape[0]
        if out is None:
            out = torch.empty((n, self.n_joints_out, 3), dtype=torch.float32, device=images.device)
        elif (tuple(out.shape) != (n, self.n_joints_out, 3) or out.dtype != torch.float32 or not out.is_contiguous()
              or out.device != images.device):
            raise ValueError(f'out must be a contiguous float32 [{n},{self.n_joints_out},3] tensor on {images.device}')
        if stream is None:
            stream = torch.cuda.current_stream(images.device).cuda_stream
        fn = self.lib.metro_infer_u8 if images.dtype == torch.uint8 else self.lib.metro_infer
        # the frozen graph's placeholder is [None,256,256,3] (main.py:109-110): any batch is accepted, in pieces of
        # at most max_batch crops (the arena's size)
        for lo in range(0, n, self.max_batch):
            cnt = min(self.max_batch, n - lo)
            _lib.check(fn(self._h, images[lo:lo + cnt].data_ptr(), cnt, out[lo:lo + cnt].data_ptr(), stream))
        return out

    def infer_coords(self, images, stream: Optional[int] = None):
        """The evaluation graph's second fetch (metro_infer_coords): returns ``(poses [n,J_out,3] mm, coords01
        [n,J_model,3])`` -- the heatmap coordinates in [0,1] that ``net_output_to_heatmap_and_coords`` returns
        (volumetric.py:234), the input of ``back_project``."""
        torch = _torch()
        if images.dim() != 4 or tuple(images.shape[1:]) != (self.spec.proc_side, self.spec.proc_side, 3) or \
                images.dtype != torch.float32 or not images.is_cuda or images.device.index != self.device:
            raise ValueError(f'expected a float32 [N,{self.spec.proc_side},{self.spec.proc_side},3] tensor on cuda:{self.device}')
        images = images.contiguous()
        n = images.shape[0]
        poses = torch.empty((n, self.n_joints_out, 3), dtype=torch.float32, device=images.device)
        coords = torch.empty((n, self.n_joints_model, 3), dtype=torch.float32, device=images.device)
        if stream is None:
            stream = torch.cuda.current_stream(images.device).cuda_stream
        for lo in range(0, n, self.max_batch):
            cnt = min(self.max_batch, n - lo)
            _lib.check(self.lib.metro_infer_coords(self._h, images[lo:lo + cnt].data_ptr(), cnt, poses[lo:lo + cnt].data_ptr(),
                                                   coords[lo:lo + cnt].data_ptr(), stream))
        return poses, coords

    def infer_host(self, images: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host buffers in / out (what ``sess.run`` does for numpy feeds, inference.py:26-27).  uint8 crops
        (the loader's format before improc.py:56-61) go through ``metro_infer_host_u8``: a quarter of the bytes."""
        if hasattr(images, 'numpy'):            # CPU torch tensor (possibly pinned): zero-copy view
            images = images.numpy()
        if images.ndim != 4 or images.shape[1:] != (self.spec.proc_side, self.spec.proc_side, 3):
            raise ValueError(f'expected [N,{self.spec.proc_side},{self.spec.proc_side},3] NHWC, got {images.shape}')
        if images.dtype not in (np.float32, np.uint8):
            raise ValueError(f'expected float32 (or uint8) images, got {images.dtype}')
        images = np.ascontiguousarray(images)
        n = images.shape[0]
        if out is None:
            out = np.empty((n, self.n_joints_out, 3), dtype=np.float32)
        elif hasattr(out, 'numpy'):
            out = out.numpy()
        if out.shape != (n, self.n_joints_out, 3) or out.dtype != np.float32 or not out.flags['C_CONTIGUOUS']:
            raise ValueError(f'out must be a C-contiguous float32 [{n},{self.n_joints_out},3] array')
        fn = self.lib.metro_infer_host_u8 if images.dtype == np.uint8 else self.lib.metro_infer_host
        for lo in range(0, n, self.max_batch):          # any batch size, like the graph's [None,...] placeholder
            cnt = min(self.max_batch, n - lo)
            _lib.check(fn(self._h, images[lo:lo + cnt].ctypes.data_as(C.c_void_p), cnt,
                          out[lo:lo + cnt].ctypes.data_as(C.c_void_p)))
        return out

    def __call__(self, images):
        if isinstance(images, np.ndarray):
            return self.infer_host(images)
        if not images.is_cuda:
            return self.infer_host(images)
        return self.infer(images)

    # -- introspection ----------------------------------------------------------------------------
    def debug_read(self, name: str) -> np.ndarray:
        n = C.c_uint64(0)
        _lib.check(self.lib.metro_debug_read(self._h, name.encode(), None, 0, C.byref(n)))
        dt = np.float16
        if self.precision != 'f16':
            dt = np.float64
        elif name == 'head' and self.head_dtype == _lib.METRO_F32:
            dt = np.float32
        buf = np.empty(n.value, dtype=dt)
        _lib.check(self.lib.metro_debug_read(self._h, name.encode(), buf.ctypes.data_as(C.c_void_p), buf.nbytes, None))
        return buf.reshape(self.max_batch, -1)

    def profile(self, images, out=None):
        torch = _torch()
        n = images.shape[0]
        if out is None:
            out = torch.empty((n, self.n_joints_out, 3), dtype=torch.float32, device=images.device)
        cnt = self.launch_count(n)
        ms = (C.c_float * (cnt + 4))()
        names = C.create_string_buffer(64 * (cnt + 4))
        k = C.c_int32(0)
        _lib.check(self.lib.metro_profile(self._h, images.data_ptr(), n, out.data_ptr(), ms, names, len(names), C.byref(k)))
        return list(zip(names.value.decode().split('\n'), list(ms)[:k.value]))

    def launch_count(self, n: int) -> int:
        k = C.c_int32(0)
        _lib.check(self.lib.metro_launch_count(self._h, n, C.byref(k)))
        return k.value

    def graph_stats(self):
        """(instantiated CUDA graphs, replays so far) of the small-batch executor (metro_graph_stats)."""
        g, r = C.c_int32(0), C.c_int64(0)
        _lib.check(self.lib.metro_graph_stats(self._h, C.byref(g), C.byref(r)))
        return g.value, r.value

    def workspace_bytes(self, n: int = 0) -> int:
        """Device bytes the handle holds for batches of up to ``n`` crops (0 = its whole arena)."""
        b = C.c_uint64(0)
        _lib.check(self.lib.metro_workspace_bytes(self._h, n, C.byref(b)))
        return b.value

    def close(self):
        if getattr(self, '_h', None):
            self.lib.metro_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_loaded = {}              # (path, device) -> MetroModel
_MAX_LOADED = 4
_DEFAULT_MAX_BATCH = 64


def estimate_pose(images, model) -> Tuple[object, np.ndarray, list]:
    """Mirror of ``estimate_pose(images_tensor, model_path)`` (inference.py:31-43): returns
    (poses [N,J,3] mm, joint_edges [E,2] int64, joint_names [J]).  ``model`` is a ``MetroModel`` or, as in the
    reference, the path of an exported ``.pb`` file (loaded once per path)."""
    if isinstance(model, (str, bytes)) and not isinstance(model, MetroModel):
        dev = images.device.index if getattr(images, 'is_cuda', False) else None
        if dev is None:
            dev = _torch().cuda.current_device()
        key = (model, dev)
        if key not in _loaded:
            while len(_loaded) >= _MAX_LOADED:          # bounded: least recently loaded model goes first
                _loaded.pop(next(iter(_loaded))).close()
            # larger batches are processed in pieces of max_batch crops (MetroModel.infer / infer_host)
            _loaded[key] = MetroModel.from_frozen_graph(model, max_batch=_DEFAULT_MAX_BATCH, device=dev)
        model = _loaded[key]
    return model(images), model.joint_edges, model.joint_names


class SoftArgmax:
    """Stand-alone heatmap decode (volumetric.py:227-235,288-306; tfu3d.py:23-25; main.py:127)."""

    def __init__(self, side: int, n_joints_model: int, stride: int, permutation: Sequence[int], depth: int = 8,
                 centered_stride: bool = True, proc_side: int = 256, box_size_mm: float = 2200.0,
                 head_dtype: str = 'f32', splits: int = 0, lanes: int = 0, word_bytes: int = 0):
        self.lib = _lib.load()
        self.perm = (C.c_int32 * len(permutation))(*permutation)
        self.n_out = len(permutation)
        self.desc = _lib.SoftargmaxDesc(side, n_joints_model, depth, stride, int(centered_stride), proc_side,
                                        box_size_mm, len(permutation), C.cast(self.perm, C.POINTER(C.c_int32)),
                                        {'f32': 0, 'f16': 1}[head_dtype], splits, lanes, word_bytes)
        self.side, self.channels = side, depth * n_joints_model
        self.head_dtype = head_dtype
        self._ws = None
        self._ws_n = -1

    def workspace_bytes(self, n: int) -> int:
        b = C.c_uint64(0)
        _lib.check(self.lib.metro_softargmax_workspace_bytes(C.byref(self.desc), n, C.byref(b)))
        return b.value

    def __call__(self, head, out=None):
        torch = _torch()
        want = torch.float32 if self.head_dtype == 'f32' else torch.float16
        if head.dim() != 4 or tuple(head.shape[1:]) != (self.side, self.side, self.channels):
            raise ValueError(f'expected NHWC [N,{self.side},{self.side},{self.channels}], got {tuple(head.shape)}')
        if head.dtype != want or not head.is_cuda:
            raise ValueError(f'expected a CUDA {want} tensor')
        head = head.contiguous()
        n = head.shape[0]
        if self._ws is None or self._ws_n != n:
            self._ws = torch.zeros(max(self.workspace_bytes(n), 16), dtype=torch.uint8, device=head.device)
            self._ws_n = n
        if out is None:
            out = torch.empty((n, self.n_out, 3), dtype=torch.float32, device=head.device)
        stream = torch.cuda.current_stream(head.device).cuda_stream
        _lib.check(self.lib.metro_softargmax(C.byref(self.desc), head.data_ptr(), n, out.data_ptr(),
                                             self._ws.data_ptr(), stream))
        return out


    def coords(self, head):
        """(poses, coords01): the decode with the heatmap coordinates in [0,1] of every model joint as a second output
        (metro_softargmax_coords; volumetric.py:234)."""
        torch = _torch()
        out = self(head)                                  # validates, sizes the workspace
        n = head.shape[0]
        c01 = torch.empty((n, self.channels // self.desc.depth, 3), dtype=torch.float32, device=head.device)
        stream = torch.cuda.current_stream(head.device).cuda_stream
        _lib.check(self.lib.metro_softargmax_coords(C.byref(self.desc), head.contiguous().data_ptr(), n, out.data_ptr(),
                                                    c01.data_ptr(), self._ws.data_ptr(), stream))
        return out, c01

    def heatmap_z(self, head):
        """``t.heatmap_pred_z`` (volumetric.py:165): depth marginal of the softmaxed heatmap, [N, J_model, D]."""
        torch = _torch()
        n = head.shape[0]
        out = torch.empty((n, self.channels // self.desc.depth, self.desc.depth), dtype=torch.float32, device=head.device)
        stream = torch.cuda.current_stream(head.device).cuda_stream
        _lib.check(self.lib.metro_heatmap_z(C.byref(self.desc), head.contiguous().data_ptr(), n, out.data_ptr(), stream))
        return out


def back_project(coords01, inv_intrinsics, z_offset, stride: int, centered_stride: bool = True, proc_side: int = 256,
                 box_size_mm: float = 2200.0):
    """The 'true-root-depth' branch of the reference's evaluation graph (src/model/volumetric.py:190-198,285) on device
    tensors: coords01 float32 ``[N, J, 3]`` (root joint last), inv_intrinsics float32 ``[N, 3, 3]``, z_offset float32
    ``[N]`` (the true root depth, or a bone-length fit's offset).  Returns absolute camera-frame ``[N, J, 3]`` mm."""
    torch = _torch()
    if coords01.dim() != 3 or coords01.shape[2] != 3 or coords01.dtype != torch.float32 or not coords01.is_cuda:
        raise ValueError('expected a CUDA float32 [N, J, 3] tensor of heatmap coordinates')
    n, j = coords01.shape[0], coords01.shape[1]
    if tuple(inv_intrinsics.shape) != (n, 3, 3) or inv_intrinsics.dtype != torch.float32 or not inv_intrinsics.is_cuda:
        raise ValueError(f'expected a CUDA float32 [{n}, 3, 3] tensor of inverse intrinsics')
    if tuple(z_offset.shape) != (n,) or z_offset.dtype != torch.float32 or not z_offset.is_cuda:
        raise ValueError(f'expected a CUDA float32 [{n}] tensor of depth offsets')
    lib = _lib.load()
    out = torch.empty_like(coords01)
    stream = torch.cuda.current_stream(coords01.device).cuda_stream
    _lib.check(lib.metro_back_project(coords01.contiguous().data_ptr(), inv_intrinsics.contiguous().data_ptr(),
                                      z_offset.contiguous().data_ptr(), n, j, stride, int(centered_stride), proc_side,
                                      box_size_mm, out.data_ptr(), stream))
    return out


def crop_homography(old_intrinsics, old_r, new_intrinsics, new_r) -> np.ndarray:
    """The homography of ``cameralib.reproject_image_fast`` (src/cameralib.py:411-413): maps a pixel of the NEW
    (crop) camera to the pixel of the OLD (frame) camera that sees the same ray; both cameras share their centre."""
    old_matrix = np.asarray(old_intrinsics, np.float64) @ np.asarray(old_r, np.float64)
    new_matrix = np.asarray(new_intrinsics, np.float64) @ np.asarray(new_r, np.float64)
    return np.linalg.solve(new_matrix.T, old_matrix.T).T.astype(np.float32)


def extract_crops(frames, homographies, side: int = 256, border_value: int = 0, out=None):
    """``reproject_image_fast`` (src/cameralib.py:406-429) on the GPU: ``frames`` is one CUDA uint8 ``[H, W, 3]`` tensor
    (every crop comes from it) or a sequence of such tensors, one per crop; ``homographies`` float32 ``[n, 3, 3]`` (host).
    Returns CUDA uint8 ``[n, side, side, 3]`` crops -- the input of ``MetroModel.infer`` -- with cv2.remap's exact
    fixed-point bilinear arithmetic."""
    torch = _torch()
    hs = np.ascontiguousarray(np.asarray(homographies, dtype=np.float32).reshape(-1, 9))
    n = hs.shape[0]
    single = hasattr(frames, 'dim')
    if not single and len(frames) != n:
        raise ValueError(f'{len(frames)} frames for {n} homographies')
    srcs = (_lib.CropSrc * max(n, 1))()
    keep = []
    for i in range(n):
        f = frames if single else frames[i]
        if f.dim() != 3 or f.shape[2] != 3 or f.dtype != torch.uint8 or not f.is_cuda or f.stride(2) != 1 or f.stride(1) != 3:
            raise ValueError('frames must be CUDA uint8 [H, W, 3] tensors with packed pixels')
        keep.append(f)
        srcs[i].frame_dev = f.data_ptr()
        srcs[i].height, srcs[i].width, srcs[i].row_stride_bytes = f.shape[0], f.shape[1], f.stride(0)
        for k in range(9):
            srcs[i].homography[k] = float(hs[i, k])
    dev = (frames if single else frames[0]).device if n else torch.device('cuda')
    if out is None:
        out = torch.empty((n, side, side, 3), dtype=torch.uint8, device=dev)
    elif tuple(out.shape) != (n, side, side, 3) or out.dtype != torch.uint8 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f'out must be a contiguous CUDA uint8 [{n},{side},{side},3] tensor')
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().metro_extract_crops(srcs, n, side, border_value, out.data_ptr(), stream))
    return out


def conv2d(x, w_hwio: np.ndarray, scale: np.ndarray, shift: np.ndarray, stride: int = 1, rate: int = 1,
           pad_lo: Optional[int] = None, relu: bool = False, out_dtype: str = 'f16', res=None, res_stride: int = 0,
           res_shift: int = 0, x2=None, w2: Optional[np.ndarray] = None, scale2=None, shift2=None):
    """Operator-level entry to the fused tcgen05 convolution (metro_conv2d); x: CUDA fp16 NHWC."""
    torch = _torch()
    lib = _lib.load()
    n, side, _, cin = x.shape
    k = w_hwio.shape[0]
    cout = w_hwio.shape[3]
    k_eff = k + (k - 1) * (rate - 1)
    if pad_lo is None:
        pad_lo = (k_eff - 1) // 2
    out_side = side if stride == 1 else side // stride
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    w_hwio, scale, shift = f32(w_hwio), f32(scale), f32(shift)
    cin2 = 0 if x2 is None else x2.shape[3]
    d = _lib.ConvDesc(n, side, cin, cout, k, stride, rate, pad_lo, int(relu), 0 if out_dtype == 'f32' else 1,
                      res_stride if res is not None else 0, res_shift, cin2)
    y = torch.empty((n, out_side, out_side, cout), dtype=torch.float32 if out_dtype == 'f32' else torch.float16,
                    device=x.device)
    y2 = None
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    if scale2 is not None:
        scale2, shift2 = f32(scale2), f32(shift2)
        y2 = torch.empty((n, out_side, out_side, cout), dtype=torch.float16, device=x.device)
    if w2 is not None:
        w2 = f32(w2)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(lib.metro_conv2d(C.byref(d), x.data_ptr(), p(w_hwio), None if x2 is None else x2.data_ptr(), p(w2),
                                p(scale), p(shift), None if res is None else res.data_ptr(), y.data_ptr(),
                                p(scale2), p(shift2), None if y2 is None else y2.data_ptr(), x.device.index or 0, stream))
    return (y, y2) if y2 is not None else y


def to_orig_cam(poses, rot_to_orig_cam, mirror_mapping: Sequence[int]):
    """``to_orig_cam`` of the reference's evaluation graph (src/model/volumetric.py:277-282) on device tensors:
    poses float32 ``[N, J, 3]``, rotations float32 ``[N, 3, 3]``; ``mirror_mapping`` as
    ``JointInfo.mirror_mapping`` (datasets.py:76-79).  Returns a new ``[N, J, 3]`` tensor."""
    torch = _torch()
    if poses.dim() != 3 or poses.shape[2] != 3 or poses.dtype != torch.float32 or not poses.is_cuda:
        raise ValueError('expected a CUDA float32 [N, J, 3] tensor of poses')
    n, j = poses.shape[0], poses.shape[1]
    if tuple(rot_to_orig_cam.shape) != (n, 3, 3) or rot_to_orig_cam.dtype != torch.float32 or not rot_to_orig_cam.is_cuda:
        raise ValueError(f'expected a CUDA float32 [{n}, 3, 3] tensor of rotations')
    if len(mirror_mapping) != j:
        raise ValueError(f'mirror_mapping has {len(mirror_mapping)} entries for {j} joints')
    lib = _lib.load()
    poses, rot = poses.contiguous(), rot_to_orig_cam.contiguous()
    out = torch.empty_like(poses)
    mm = (C.c_int32 * j)(*[int(v) for v in mirror_mapping])
    stream = torch.cuda.current_stream(poses.device).cuda_stream
    _lib.check(lib.metro_to_orig_cam(poses.data_ptr(), rot.data_ptr(), C.cast(mm, C.c_void_p), n, j, out.data_ptr(), stream))
    return out
