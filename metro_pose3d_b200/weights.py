"""Synthetic weights for the MeTRo graph and their canonical flat serialisation.

Variable naming follows the TF-slim scopes the reference creates under ``MainPart/resnet_v2_*``
(src/model/architectures.py:24, src/model/resnet_v2.py:117-138,219-236), so that a future
frozen-graph importer can fill the same dictionary.

The flat blob handed to ``metro_create`` is the concatenation, as float32, of every array in
``blob_order(spec)``; conv filters are in TF's HWIO layout.  csrc/plan.cpp walks the same order.

Initialisation (there are no published weights on disk and no network): filters use slim's
``variance_scaling_initializer()`` (src/model/architectures.py:16: factor 2, fan-in, truncated
normal), biases N(0, 0.01^2), BN gamma~U(.5,1.5), beta~N(0,.1^2), mean~N(0,.1^2), var~U(.5,1.5) so
BN-folding mistakes are visible.  ``residual_gain`` scales every conv3 filter: with fixed (not
batch-estimated) BN statistics an undamped random pre-activation ResNet doubles its variance per
unit (2^33 for ResNet-101), which no trained network does and which would overflow the fp16
activations the reference computes in (src/options.py:73).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np

from .spec import NetSpec, Conv

_TRUNC_STD_CORRECTION = 0.87962566103423978  # std of N(0,1) truncated to [-2, 2]


def _bn_names(prefix: str) -> List[str]:
    return [f'{prefix}/{p}' for p in ('gamma', 'beta', 'moving_mean', 'moving_variance')]


def blob_order(spec: NetSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """[(variable name, shape)] in serialisation order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(c: Conv, scope: str):
        out.append((f'{scope}/weights', (c.k, c.k, c.cin, c.cout)))
        if c.has_bias:
            out.append((f'{scope}/biases', (c.cout,)))
        if c.has_bn:
            out.extend((n, (c.cout,)) for n in _bn_names(f'{scope}/BatchNorm'))

    conv(spec.root, 'conv1')
    for u in spec.units:
        s = f'{u.name}/bottleneck_v2'
        out.extend((n, (u.cin,)) for n in _bn_names(f'{s}/preact'))
        if u.shortcut is not None:
            conv(u.shortcut, f'{s}/shortcut')
        conv(u.conv1, f'{s}/conv1')
        conv(u.conv2, f'{s}/conv2')
        conv(u.conv3, f'{s}/conv3')
    out.extend((n, (spec.feat_channels,)) for n in _bn_names('postnorm'))
    conv(spec.logits, 'logits')
    return out


def blob_size(spec: NetSpec) -> int:
    return sum(int(np.prod(s)) for _, s in blob_order(spec))


def synth_weights(spec: NetSpec, seed: int = 0, residual_gain: float = 0.25,
                  logit_gain: float = 1.0) -> 'OrderedDict[str, np.ndarray]':
    w: 'OrderedDict[str, np.ndarray]' = OrderedDict()
    for idx, (name, shape) in enumerate(blob_order(spec)):
        rng = np.random.default_rng([seed, idx])
        leaf = name.rsplit('/', 1)[1]
        if leaf == 'weights':
            fan_in = shape[0] * shape[1] * shape[2]
            std = np.sqrt(2.0 / fan_in) / _TRUNC_STD_CORRECTION
            a = rng.standard_normal(shape)
            bad = np.abs(a) > 2.0
            while bad.any():                       # truncated normal by resampling
                a[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(a) > 2.0
            a *= std
            if name.endswith('conv3/weights'):
                a *= residual_gain
            if name == 'logits/weights':
                a *= logit_gain
        elif leaf == 'biases':
            a = 0.01 * rng.standard_normal(shape)
        elif leaf == 'gamma':
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf in ('beta', 'moving_mean'):
            a = 0.1 * rng.standard_normal(shape)
        elif leaf == 'moving_variance':
            a = rng.uniform(0.5, 1.5, shape)
        else:
            raise AssertionError(name)
        w[name] = a.astype(np.float32)
    return w


def pack_blob(spec: NetSpec, weights: Dict[str, np.ndarray]) -> np.ndarray:
    parts = []
    for name, shape in blob_order(spec):
        a = np.asarray(weights[name], dtype=np.float32)
        if a.shape != tuple(shape):
            raise ValueError(f'{name}: expected shape {shape}, got {a.shape}')
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def unpack_blob(spec: NetSpec, blob: np.ndarray) -> 'OrderedDict[str, np.ndarray]':
    blob = np.asarray(blob, dtype=np.float32).reshape(-1)
    if blob.size != blob_size(spec):
        raise ValueError(f'blob has {blob.size} floats, spec needs {blob_size(spec)}')
    w: 'OrderedDict[str, np.ndarray]' = OrderedDict()
    off = 0
    for name, shape in blob_order(spec):
        n = int(np.prod(shape))
        w[name] = blob[off:off + n].reshape(shape)
        off += n
    return w


def synth_images(n: int, seed: int = 1000, side: int = 256) -> np.ndarray:
    """float32 NHWC in [0,1) (input contract: inference.py:17-18, improc.py:56-61)."""
    rng = np.random.default_rng(seed)
    return rng.random((n, side, side, 3), dtype=np.float32)


def synth_head(n: int, side: int, n_joints: int, depth: int = 8, seed: int = 0,
               sigma: float = 3.0, peak: float = 10.0) -> np.ndarray:
    """Stand-alone soft-argmax input: NHWC float32 ~N(0, sigma^2) with one planted
    +peak per (n, j); channel order is depth-major c = d*J + j (volumetric.py:231-232)."""
    rng = np.random.default_rng(seed)
    c = depth * n_joints
    x = (sigma * rng.standard_normal((n, side, side, c), dtype=np.float32))
    hh = rng.integers(0, side, (n, n_joints))
    ww = rng.integers(0, side, (n, n_joints))
    dd = rng.integers(0, depth, (n, n_joints))
    ni = np.arange(n)[:, None]
    ji = np.arange(n_joints)[None, :]
    x[ni, hh, ww, dd * n_joints + ji] += np.float32(peak)
    return x
