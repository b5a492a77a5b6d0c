"""CPU oracle: a restatement of the reference's exported inference graph.  TEST INFRASTRUCTURE.

    PIN STATUS: pinned at graph level against the reference's own code run here; the TensorFlow 1.13.1
    op kernels themselves are unpinned.  The reference holds no tests, golden vectors, fixtures or weights
     and its own implementation needs TensorFlow 1.13.1 + tf.contrib.slim,
    which cannot be installed here (no wheel, no network, no py3.12 build).  oracle/gen_golden.py
    therefore imports the reference's Python files from /root/reference/src and executes them eagerly in
    float64 over oracle/tf_shim.py; the outputs are committed under tests/golden/ and this module is
    checked against them (tests/test_golden.py).  Every graph-level decision is thus the reference's;
    the primitive op semantics (SAME/VALID padding, HWIO filters, FusedBatchNorm inference form,
    reduce_max/exp/sum softmax, linspace grids) are restated from TensorFlow's documentation at the
    reference's call sites, cited line by line below, and additionally checked against self-derived
    analytic known answers (tests/test_oracle_*.py) and an independent scalar numpy convolution
    (tests/test_oracle_backbone.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (metro_pose3d_b200/) never does.

Precision modes
    'fp64'  every tensor float64 -- ground truth.
    'fp32'  every tensor float32 -- what a TF-CPU float32 graph computes (up to summation order).
    'half'  operands rounded to float16 at exactly the points where the CUDA path stores float16
            (DESIGN.md "rounding points"), products/accumulation in float64 -- the tight
            per-layer check for the tensor-core kernels.  The reference's default backbone dtype
            is float16 (src/options.py:73, src/model/architectures.py:29).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from metro_pose3d_b200.spec import NetSpec, Conv, Unit, BN_EPS, BOX_SIZE_MM


# =============================================================================================
# Decode: head output -> [N, J_out, 3] root-relative millimetres
# =============================================================================================
def softmax_ref(x: np.ndarray, axis) -> np.ndarray:
    """src/tfu.py:466-471."""
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def decode_heatmap_ref(inp: np.ndarray, axes: Sequence[int]):
    """src/tfu.py:474-499 (multiple-axes branch): marginal over the other heatmap axes, then the
    expectation of linspace(0, 1, n) along the axis."""
    res = []
    for ax in axes:
        others = tuple(a for a in axes if a != ax)
        marg = np.sum(inp, axis=others, keepdims=True)
        shape = [1] * inp.ndim
        shape[ax] = inp.shape[ax]
        grid = np.linspace(0.0, 1.0, inp.shape[ax]).astype(inp.dtype).reshape(shape)
        dec = np.sum(grid * marg, axis=ax, keepdims=True)
        res.append(np.squeeze(dec, axis=tuple(axes)))
    return res


def decode_ref(head_nhwc: np.ndarray, n_joints: int, stride: int, permutation: Sequence[int],
               centered_stride: bool = True, proc_side: int = 256,
               box_size_mm: float = BOX_SIZE_MM, dtype=np.float64,
               return_coords01: bool = False) -> np.ndarray:
    """head [N, H, W, D*J] (what `architectures.resnet` returns, here NHWC) -> poses [N, J_out, 3].

    Follows, in order:
      volumetric.py:227-235  std_to_nchw -> reshape [N,D,J,H,W] -> transpose [N,J,H,W,D]
                             -> softmax over (H,W,D) -> decode axes [3,2,4] (x<-W, y<-H, z<-D)
      volumetric.py:288-306  heatmap_to_image / heatmap_to_metric
      tfu3d.py:23-25         root_relative: subtract the LAST model joint
      main.py:119-127        tf.gather(permutation, axis=1)
    """
    x = np.asarray(head_nhwc).astype(dtype)
    n, h, w, c = x.shape
    depth = c // n_joints
    nchw = np.transpose(x, (0, 3, 1, 2))                              # tfu.std_to_nchw
    reshaped = nchw.reshape(n, depth, n_joints, h, w)                 # volumetric.py:231
    transposed = np.transpose(reshaped, (0, 2, 3, 4, 1))              # volumetric.py:232
    softmaxed = softmax_ref(transposed, axis=(2, 3, 4))               # volumetric.py:233
    coords = np.stack(decode_heatmap_ref(softmaxed, [3, 2, 4]), axis=-1)   # volumetric.py:234
    if return_coords01:
        return coords
    last_image_pixel = proc_side - 1                                  # volumetric.py:290
    last_receptive_center = last_image_pixel - (last_image_pixel % stride) - 1
    c2d = coords[..., :2] * dtype(last_receptive_center)
    if centered_stride:
        c2d = c2d + stride // 2
    c2d = c2d * dtype(box_size_mm) / proc_side                        # volumetric.py:304-305
    metric = np.concatenate([c2d, coords[..., 2:] * dtype(box_size_mm)], axis=-1)
    rootrel = metric - metric[:, -1:, :]                              # tfu3d.py:23-25
    return rootrel[:, list(permutation), :]                           # main.py:127


# =============================================================================================
# Backbone
# =============================================================================================
_TORCH_DT = {'fp64': torch.float64, 'fp32': torch.float32, 'half': torch.float64}


class OracleNet:
    """`architectures.resnet` + decode on CPU with torch ops.  Activations are NCHW internally
    (the reference's default data_format, src/options.py:92); I/O is NHWC like the graph's."""

    def __init__(self, spec: NetSpec, weights: Dict[str, np.ndarray], permutation: Sequence[int],
                 mode: str = 'fp64'):
        assert mode in _TORCH_DT
        self.spec, self.mode, self.perm = spec, mode, list(permutation)
        self.dt = _TORCH_DT[mode]
        self.w = weights
        self.trace: Optional[Dict[str, np.ndarray]] = None   # per-layer NHWC dumps when enabled

    # ---- helpers ----------------------------------------------------------------------------
    def _q(self, t: torch.Tensor) -> torch.Tensor:
        """Storage rounding point: float16 in 'half' mode, identity otherwise."""
        if self.mode == 'half':
            return t.to(torch.float16).to(self.dt)
        return t

    def _filter(self, name: str) -> torch.Tensor:
        w = torch.from_numpy(np.ascontiguousarray(self.w[name + '/weights']))   # HWIO
        if self.mode == 'half':
            w = w.to(torch.float16)
        return w.to(self.dt).permute(3, 2, 0, 1).contiguous()                   # -> OIHW

    def _vec(self, name: str) -> torch.Tensor:
        # per-channel vectors stay float32 on the device; the oracle keeps them at >= that
        v = torch.from_numpy(np.ascontiguousarray(self.w[name]))
        return v.to(torch.float32 if self.mode == 'fp32' else torch.float64)

    def _bn_affine(self, scope: str):
        """FusedBatchNorm inference form gamma*(x-mean)/sqrt(var+eps)+beta as scale/shift
        (architectures.py:9-11: epsilon=1e-5, scale=True)."""
        g, b = self._vec(scope + '/gamma'), self._vec(scope + '/beta')
        m, v = self._vec(scope + '/moving_mean'), self._vec(scope + '/moving_variance')
        if self.mode == 'half':
            # the device holds scale/shift as float32 computed in double on the host
            scale = (g / torch.sqrt(v + BN_EPS)).to(torch.float32).to(torch.float64)
            shift = (b - m * (g / torch.sqrt(v + BN_EPS))).to(torch.float32).to(torch.float64)
        else:
            scale = g / torch.sqrt(v + BN_EPS)
            shift = b - m * scale
        return scale.to(self.dt).view(1, -1, 1, 1), shift.to(self.dt).view(1, -1, 1, 1)

    def _conv(self, x: torch.Tensor, c: Conv, scope: str) -> torch.Tensor:
        """conv2d_same (resnet_utils.py:82-135): explicit zero pad (pad_lo, pad_hi) then VALID;
        for stride 1 / centred stride the pads are TF's SAME pads (spec.same_pad)."""
        if c.pad_lo or c.pad_hi:
            x = F.pad(x, (c.pad_lo, c.pad_hi, c.pad_lo, c.pad_hi))
        return F.conv2d(x, self._filter(scope), None, stride=c.stride, dilation=c.rate)

    def _dump(self, name: str, t: torch.Tensor):
        if self.trace is not None:
            self.trace[name] = t.permute(0, 2, 3, 1).contiguous().numpy().copy()

    # ---- graph ------------------------------------------------------------------------------
    def _unit(self, x: torch.Tensor, pre: torch.Tensor, u: Unit, next_bn: str):
        """bottleneck (resnet_v2.py:84-139).  `x` is the raw unit input, `pre` = relu(bn(x))
        (:119).  Returns (raw output, relu(next_bn(raw output)))."""
        s = f'{u.name}/bottleneck_v2'
        sh = u.shift
        # conv1 -> BN -> ReLU (:127-128)
        sc1, sf1 = self._bn_affine(s + '/conv1/BatchNorm')
        r = self._q(F.relu(self._conv(pre, u.conv1, s + '/conv1') * sc1 + sf1))
        self._dump(u.name + '/conv1', r)
        # conv2 3x3 (stride, rate, centred) -> BN -> ReLU (:130-132)
        sc2, sf2 = self._bn_affine(s + '/conv2/BatchNorm')
        r = self._q(F.relu(self._conv(r, u.conv2, s + '/conv2') * sc2 + sf2))
        self._dump(u.name + '/conv2', r)
        # conv3 1x1 + bias (:134-136)
        r = self._conv(r, u.conv3, s + '/conv3') + self._vec(s + '/conv3/biases').to(self.dt).view(1, -1, 1, 1)
        # shortcut (:120-125)
        if u.shortcut is None:
            shortcut = x[:, :, sh::u.stride, sh::u.stride]          # subsample(_shift(inputs))
        else:
            xs = pre[:, :, sh:, sh:] if sh else pre
            shortcut = (F.conv2d(xs, self._filter(s + '/shortcut'), None, stride=u.stride)
                        + self._vec(s + '/shortcut/biases').to(self.dt).view(1, -1, 1, 1))
        # 'half': the device fuses shortcut + conv3 + biases in one float32 accumulator and rounds
        # once; the float64 sum here is that accumulator without its rounding noise.
        out = self._q(shortcut + r)                                    # :138
        self._dump(u.name + '/out', out)
        scn, sfn = self._bn_affine(next_bn)
        pre_next = self._q(F.relu(out * scn + sfn))
        return out, pre_next

    def forward_head(self, images_nhwc: np.ndarray) -> np.ndarray:
        """`architectures.resnet` (architectures.py:24-35): NHWC [N,256,256,3] in [0,1] ->
        NHWC [N,Hh,Hh,D*J]."""
        sp = self.spec
        x = torch.from_numpy(np.ascontiguousarray(images_nhwc)).to(self.dt).permute(0, 3, 1, 2)
        x = self._q(x)                                                 # architectures.py:29 cast
        # conv1 7x7/2 + bias, no BN / activation (resnet_v2.py:219-220)
        x = self._q(self._conv(x, sp.root, 'conv1')
                    + self._vec('conv1/biases').to(self.dt).view(1, -1, 1, 1))
        self._dump('conv1', x)
        # pool1: zero pad (1,1) + 3x3/2 VALID max-pool (resnet_utils.py:177-185) -- zeros, not -inf
        x = F.max_pool2d(F.pad(x, (1, 1, 1, 1)), 3, 2)
        self._dump('pool1', x)
        names = [f'{u.name}/bottleneck_v2/preact' for u in sp.units] + ['postnorm']
        sc, sf = self._bn_affine(names[0])
        pre = self._q(F.relu(x * sc + sf))
        for i, u in enumerate(sp.units):
            x, pre = self._unit(x, pre, u, names[i + 1])
        self._dump('postnorm', pre)                                    # resnet_v2.py:229
        # logits 1x1 + bias (:234-236), cast to float32 (architectures.py:34)
        y = (self._conv(pre, sp.logits, 'logits')
             + self._vec('logits/biases').to(self.dt).view(1, -1, 1, 1))
        if self.mode in ('half', 'fp32'):
            y = y.to(torch.float32)
        return y.permute(0, 2, 3, 1).contiguous().numpy()

    def decode(self, head_nhwc: np.ndarray) -> np.ndarray:
        dt = np.float32 if self.mode == 'fp32' else np.float64
        return decode_ref(head_nhwc, self.spec.n_joints, self.spec.stride, self.perm,
                          self.spec.centered_stride, self.spec.proc_side, dtype=dt)

    def __call__(self, images_nhwc: np.ndarray) -> np.ndarray:
        return self.decode(self.forward_head(images_nhwc))


def estimate_pose_oracle(images_nhwc, spec, weights, dataset, mode='fp32'):
    """The reference contract (inference.py:31-43): poses, edges, names."""
    from metro_pose3d_b200.joints import exported_joint_info, export_permutation
    ji = exported_joint_info(dataset)
    poses = OracleNet(spec, weights, export_permutation(dataset), mode)(images_nhwc)
    return poses, np.asarray(ji.edges, dtype=np.int64), list(ji.names)


# =============================================================================================
# Single fused convolution (operator-level oracle for metro_conv2d)
# =============================================================================================
def conv2d_fused_ref(x_nhwc, w_hwio, scale, shift, stride=1, rate=1, pad_lo=None, pad_hi=None, relu=False,
                     res_nhwc=None, res_stride=0, res_shift=0, x2_nhwc=None, w2=None, scale2=None, shift2=None,
                     out_f16=True):
    """float64 evaluation of what one launch of the tcgen05 kernel computes, from the same fp16
    operands: y = conv(x, fp16(w)) [+ conv1x1(x2, fp16(w2))] * scale + shift [+ res[:, s::r, s::r]],
    ReLU?, rounded to fp16 (or fp32); y2 = fp16(relu(float(y) * scale2 + shift2)).
    Padding follows conv2d_same (resnet_utils.py:82-135)."""
    k = w_hwio.shape[0]
    k_eff = k + (k - 1) * (rate - 1)
    if pad_lo is None:
        pad_lo = (k_eff - 1) // 2
    if pad_hi is None:
        pad_hi = (k_eff - 1) - pad_lo
    x = torch.from_numpy(np.asarray(x_nhwc, dtype=np.float64)).permute(0, 3, 1, 2)
    w = torch.from_numpy(np.asarray(w_hwio, dtype=np.float32)).to(torch.float16).to(torch.float64)
    y = F.conv2d(F.pad(x, (pad_lo, pad_hi, pad_lo, pad_hi)), w.permute(3, 2, 0, 1).contiguous(), None,
                 stride=stride, dilation=rate)
    if x2_nhwc is not None:
        x2 = torch.from_numpy(np.asarray(x2_nhwc, dtype=np.float64)).permute(0, 3, 1, 2)
        ww = torch.from_numpy(np.asarray(w2, dtype=np.float32)).to(torch.float16).to(torch.float64)
        y = y + F.conv2d(x2, ww.permute(3, 2, 0, 1).contiguous())
    sc = torch.from_numpy(np.asarray(scale, dtype=np.float32)).to(torch.float64).view(1, -1, 1, 1)
    sf = torch.from_numpy(np.asarray(shift, dtype=np.float32)).to(torch.float64).view(1, -1, 1, 1)
    y = y * sc + sf
    if res_nhwc is not None and res_stride > 0:
        r = torch.from_numpy(np.asarray(res_nhwc, dtype=np.float64)).permute(0, 3, 1, 2)
        y = y + r[:, :, res_shift::res_stride, res_shift::res_stride]
    if relu:
        y = F.relu(y)
    y_exact = y.permute(0, 2, 3, 1).contiguous().numpy()
    yq = y.to(torch.float16 if out_f16 else torch.float32).to(torch.float64)
    y2 = None
    if scale2 is not None:
        s2 = torch.from_numpy(np.asarray(scale2, dtype=np.float32)).to(torch.float64).view(1, -1, 1, 1)
        f2 = torch.from_numpy(np.asarray(shift2, dtype=np.float32)).to(torch.float64).view(1, -1, 1, 1)
        y2 = F.relu(yq * s2 + f2).permute(0, 2, 3, 1).contiguous().numpy()
    return y_exact, y2


def conv2d_naive(x_nhwc, w_hwio, stride, rate, pad_lo, pad_hi):
    """Independent scalar-loop convolution (numpy only) used to pin the torch-based oracle:
    out[n,oh,ow,co] = sum_{kh,kw,ci} xpad[n, oh*s + kh*r, ow*s + kw*r, ci] * w[kh,kw,ci,co]."""
    x = np.asarray(x_nhwc, dtype=np.float64)
    w = np.asarray(w_hwio, dtype=np.float64)
    n, h, wd, ci = x.shape
    k = w.shape[0]
    xp = np.zeros((n, h + pad_lo + pad_hi, wd + pad_lo + pad_hi, ci))
    xp[:, pad_lo:pad_lo + h, pad_lo:pad_lo + wd] = x
    k_eff = k + (k - 1) * (rate - 1)
    oh = (xp.shape[1] - k_eff) // stride + 1
    ow = (xp.shape[2] - k_eff) // stride + 1
    out = np.zeros((n, oh, ow, w.shape[3]))
    for i in range(oh):
        for j in range(ow):
            for kh in range(k):
                for kw in range(k):
                    out[:, i, j, :] += xp[:, i * stride + kh * rate, j * stride + kw * rate, :] @ w[kh, kw]
    return out


def to_orig_cam_ref(poses: np.ndarray, rot_to_orig_cam: np.ndarray, mirror_mapping: Sequence[int]) -> np.ndarray:
    """Post-path step of the reference's evaluation graph (src/model/volumetric.py:277-282):
    x' = einsum('Bij,BCj->BCi', R, x) (matmul_joint_coords, :221-222); where det(R) > 0 keep x', otherwise --
    the crop was flipped horizontally, data_loading.py:80-83 -- take x' with left and right joints swapped
    (gather by JointInfo.mirror_mapping, datasets.py:76-79).  float64."""
    x = np.asarray(poses, np.float64)
    r = np.asarray(rot_to_orig_cam, np.float64)
    y = np.einsum('bij,bcj->bci', r, x)
    keep = np.linalg.det(r) > 0
    return np.where(keep[:, None, None], y, y[:, list(mirror_mapping)])


def true_root_depth_ref(coords01: np.ndarray, inv_intrinsics: np.ndarray, root_z: np.ndarray, stride: int,
                        centered_stride: bool = True, proc_side: int = 256, box_size_mm: float = BOX_SIZE_MM) -> np.ndarray:
    """Absolute-scale variant 'true-root-depth' of the evaluation graph (src/model/volumetric.py:190-198): the 2D part of
    the heatmap coordinates goes to image pixels (heatmap_to_image, :288-295), to homogeneous coordinates (:225-226) and
    through the inverse intrinsics (matmul_joint_coords, :221-222); the depth relative to the root joint (the LAST one)
    is scaled to millimetres; back_project (:285) multiplies the rays by (relative depth + true root depth).
    Returns absolute camera-frame coordinates [N, J, 3] (float64)."""
    c = np.asarray(coords01, np.float64)
    last = proc_side - 1
    lrc = last - (last % stride) - 1
    im = c[..., :2] * lrc + (stride // 2 if centered_stride else 0)
    homog = np.concatenate([im, np.ones_like(im[..., :1])], axis=-1)
    rays = np.einsum('bij,bcj->bci', np.asarray(inv_intrinsics, np.float64), homog)
    dz = (c[..., 2] - c[:, -1:, 2]) * box_size_mm
    return rays * (dz + np.asarray(root_z, np.float64)[:, None])[..., None]


def heatmap_pred_z_ref(head_nhwc: np.ndarray, n_joints: int) -> np.ndarray:
    """t.heatmap_pred_z (src/model/volumetric.py:165): the softmax over (H, W, D) summed over H and W -> [N, J, D]."""
    x = np.asarray(head_nhwc, np.float64)
    n, h, w, c = x.shape
    d = c // n_joints
    t = x.reshape(n, h, w, d, n_joints).transpose(0, 4, 1, 2, 3)            # [N, J, H, W, D]
    return softmax_ref(t, axis=(2, 3, 4)).sum(axis=(2, 3))
