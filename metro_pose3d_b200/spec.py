"""Layer plan of the MeTRo inference graph (pure Python, no compute).

``NetSpec.expand()`` replays the reference's graph builders and returns the exact list of
convolutions / pooling / residual units that the exported ``.pb`` contains:

* block table + centred-stride selection   -- src/model/resnet_v2.py:272-312
* root conv, pool1, postnorm, logits       -- src/model/resnet_v2.py:209-236
* stride / atrous-rate bookkeeping         -- src/model/resnet_utils.py:307-350
* per-unit structure (pre-activation)      -- src/model/resnet_v2.py:84-139
* padding conventions                      -- src/model/resnet_utils.py:82-185

The same replay exists in C++ (csrc/plan.cpp) for the C-ABI library; ``tests/test_plan_abi.py``
checks that both produce the same table so the oracle and the CUDA path cannot drift.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

PROC_SIDE = 256        # src/options.py:41
DEPTH = 8              # src/options.py:113
BOX_SIZE_MM = 2200.0   # src/options.py:119
BN_EPS = 1e-5          # src/model/architectures.py:10

_UNITS = {'resnet_v2_50': (3, 4, 6, 3), 'resnet_v2_101': (3, 4, 23, 3)}
_BASE_DEPTH = (64, 128, 256, 512)
_BLOCK_STRIDE = (2, 2, 2, 1)


def same_pad(n: int, k_eff: int, stride: int):
    """TensorFlow 'SAME' padding: returns (out, pad_lo, pad_hi)."""
    out = -(-n // stride)
    total = max((out - 1) * stride + k_eff - n, 0)
    return out, total // 2, total - total // 2


@dataclass
class Conv:
    name: str
    cin: int
    cout: int
    k: int
    stride: int
    rate: int
    pad_lo: int           # zero rows/cols added before the first input row/col
    pad_hi: int
    in_side: int
    out_side: int
    has_bias: bool        # slim adds a bias iff normalizer_fn is None (Q7)
    has_bn: bool          # conv -> BN -> ReLU (bottleneck conv1 / conv2)
    relu: bool

    @property
    def flops(self) -> int:
        return 2 * self.out_side * self.out_side * self.cout * self.cin * self.k * self.k

    @property
    def n_params(self) -> int:
        return self.k * self.k * self.cin * self.cout


@dataclass
class Unit:
    name: str             # 'block2/unit_4'
    cin: int
    depth: int
    cb: int
    stride: int           # stride actually applied (1 once the atrous regime starts)
    rate: int
    shift: int            # 1: centred stride -> shortcut & conv2 sample odd pixels (Q5)
    in_side: int
    out_side: int
    shortcut: Optional[Conv]  # None => identity (sub-sampled by `stride`, offset `shift`)
    conv1: Conv = None
    conv2: Conv = None
    conv3: Conv = None


@dataclass
class NetSpec:
    arch: str = 'resnet_v2_50'
    stride: int = 16
    n_joints: int = 17        # joints predicted by the head (J_model)
    depth: int = DEPTH
    centered_stride: bool = True
    proc_side: int = PROC_SIDE

    root: Conv = field(init=False, default=None)
    units: List[Unit] = field(init=False, default_factory=list)
    logits: Conv = field(init=False, default=None)

    def __post_init__(self):
        self.expand()

    # ------------------------------------------------------------------------------------------
    def expand(self):
        if self.arch not in _UNITS:
            raise ValueError(f'unknown architecture {self.arch!r}')
        if self.stride % 4 != 0:
            # resnet_v2.py:213-214
            raise ValueError('The output_stride needs to be a multiple of 4.')
        target = self.stride // 4  # resnet_v2.py:215 (true division; exact for multiples of 4)

        # centred-stride block selection, resnet_v2.py:277-281 (rn50) / :299-302 (rn101, no guard)
        c = [False, False, False]
        if self.centered_stride:
            if self.arch == 'resnet_v2_50':
                i_last = int(round(math.log2(self.stride))) - 3
                if i_last >= 0:
                    if i_last > 2:
                        raise ValueError('The target output_stride cannot be reached.')
                    c[i_last] = True
            else:
                i_last = int(math.log2(self.stride)) - 3
                if i_last > 2:
                    raise ValueError('The target output_stride cannot be reached.')
                c[i_last] = True  # i_last == -1 wraps to block3 exactly as the Python list does

        side = self.proc_side
        # conv1: conv2d_same(64, 7, stride 2) non-centred => explicit pad (3,3) + VALID (Q4)
        out = (side + 6 - 7) // 2 + 1
        self.root = Conv('conv1', 3, 64, 7, 2, 1, 3, 3, side, out, True, False, False)
        side = out
        # pool1: max_pool2d_same(3, stride 2), never centred, zero padded (Q6)
        self.pool_in = side
        side = (side + 2 - 3) // 2 + 1
        self.pool_out = side

        self.units = []
        current_stride, rate, cin = 1, 1, 64
        for b, n_units in enumerate(_UNITS[self.arch]):
            cb = _BASE_DEPTH[b]
            depth = 4 * cb
            for u in range(n_units):
                unit_stride = _BLOCK_STRIDE[b] if u == n_units - 1 else 1
                unit_centered = c[b] if (u == n_units - 1 and b < 3) else False
                # resnet_utils.py:325-333
                if current_stride == target:
                    s, r = 1, rate
                    rate *= unit_stride
                else:
                    s, r = unit_stride, 1
                    current_stride *= unit_stride
                    if current_stride > target:
                        raise ValueError('The target output_stride cannot be reached.')
                shift = 1 if (unit_centered and s == 2) else 0
                name = f'block{b + 1}/unit_{u + 1}'
                k_eff = 3 + 2 * (r - 1)
                if s == 1 or unit_centered:
                    o, lo, hi = same_pad(side, k_eff, s)          # resnet_utils.py:120-123
                else:
                    lo = (k_eff - 1) // 2                          # resnet_utils.py:125-135
                    hi = (k_eff - 1) - lo
                    o = (side + lo + hi - k_eff) // s + 1
                sc = None
                if depth != cin:
                    # projection shortcut on the pre-activation, bias, no BN (resnet_v2.py:123-125)
                    sc = Conv(name + '/shortcut', cin, depth, 1, s, 1, 0, 0, side, o,
                              True, False, False)
                unit = Unit(name, cin, depth, cb, s, r, shift, side, o, sc)
                unit.conv1 = Conv(name + '/conv1', cin, cb, 1, 1, 1, 0, 0, side, side,
                                  False, True, True)
                unit.conv2 = Conv(name + '/conv2', cb, cb, 3, s, r, lo, hi, side, o,
                                  False, True, True)
                unit.conv3 = Conv(name + '/conv3', cb, depth, 1, 1, 1, 0, 0, o, o,
                                  True, False, False)
                self.units.append(unit)
                side, cin = o, depth
        if current_stride != target:
            raise ValueError('The target output_stride cannot be reached.')
        self.feat_side = side
        self.feat_channels = cin
        self.logits = Conv('logits', cin, self.depth * self.n_joints, 1, 1, 1, 0, 0, side, side,
                           True, False, False)
        return self

    # ------------------------------------------------------------------------------------------
    @property
    def convs(self) -> List[Conv]:
        out = [self.root]
        for u in self.units:
            if u.shortcut is not None:
                out.append(u.shortcut)
            out += [u.conv1, u.conv2, u.conv3]
        out.append(self.logits)
        return out

    @property
    def flops_per_crop(self) -> int:
        return sum(c.flops for c in self.convs)

    @property
    def n_conv_params(self) -> int:
        return sum(c.n_params for c in self.convs)

    @property
    def head_channels(self) -> int:
        return self.depth * self.n_joints

    @property
    def last_receptive_center(self) -> int:
        # volumetric.py:288-291
        last = self.proc_side - 1
        return last - (last % self.stride) - 1

    def softargmax_bytes_per_crop(self, j_out: int, head_itemsize: int = 4) -> int:
        return self.feat_side ** 2 * self.head_channels * head_itemsize + j_out * 3 * 4


CONFIGS = {
    # BASELINE.json configs (name -> (arch, stride, joints-set, batch, gpus))
    'A': ('resnet_v2_50', 32, 'h36m', 1, 0),
    'B': ('resnet_v2_50', 16, 'h36m', 256, 1),
    'C': ('resnet_v2_50', 8, 'coco19', 512, 8),
    'D': ('resnet_v2_101', 16, 'coco19', 256, 1),
    'E': ('resnet_v2_101', 4, 'coco19', 1024, 8),
}
