"""TensorFlow-free reader of the reference's exported model.

``main.export()`` (src/main.py:106-160) freezes the inference graph to a binary ``GraphDef``:
``convert_variables_to_constants`` turns every variable under ``MainPart/resnet_v2_{50,101}/``
(src/model/architectures.py:24) into a ``Const`` node, ``TransformGraph`` then runs
merge_duplicate_nodes / strip_unused_nodes / fold_constants / fold_batch_norms, and the three fetches are
``output`` (a GatherV2 with the export permutation), ``joint_names`` and ``joint_edges``.

This module parses that file with a ~100-line protobuf wire-format reader (field numbers from
tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types}.proto, TF 1.13) and
recovers, by walking the graph rather than trusting constant names (fold_constants renames them):

* per ``Conv2D`` scope: the HWIO filter, the ``BiasAdd`` vector, the four ``FusedBatchNorm`` vectors -- or, for a
  batch norm built from primitive ops and folded by ``fold_constants`` / ``fold_batch_norms``, the remaining
  ``batchnorm/mul_1`` scale and ``batchnorm/add_1`` offset (the scale already inside the filter behind a
  convolution), re-expressed as gamma / beta / mean 0 / variance 1 - eps;
* the architecture (scope prefix), the head width (``depth * n_joints``), the stride (product of the conv
  strides), the export permutation, ``joint_names`` and ``joint_edges``.

The result is the same ``{name: float32 array}`` dictionary ``weights.blob_order`` serialises for
``metro_create``.  fp16 constants (the reference's default FLAGS.dtype) are widened to float32 exactly.
No TensorFlow-produced file is available offline, so the tests round-trip through
``tests/pb_writer.py`` (a minimal encoder of the same messages); the parser itself follows the public
proto definitions only.
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

from .spec import BN_EPS, NetSpec
from .weights import blob_order

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_INT32, DT_STRING, DT_INT64, DT_HALF = 1, 3, 7, 9, 19
_PASS_THROUGH = ('Identity', 'Cast', 'StopGradient', 'Snapshot')


# ---------------------------------------------------------------------------------------------
# protobuf wire format
# ---------------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yields (field_number, wire_type, value) of one message; length-delimited values are memoryviews."""
    pos, end = 0, len(buf)
    view = memoryview(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(view[pos:pos + 8]); pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val = view[pos:pos + n]; pos += n
        elif wt == 5:
            val = bytes(view[pos:pos + 4]); pos += 4
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        yield num, wt, val


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val) -> List[int]:
    buf, pos, out = bytes(val), 0, []
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(_signed(v))
    return out


# ---------------------------------------------------------------------------------------------
# TensorProto / AttrValue / NodeDef / GraphDef
# ---------------------------------------------------------------------------------------------
def _parse_shape(buf) -> List[int]:
    dims = []
    for num, _, val in _fields(bytes(buf)):
        if num == 2:                                   # TensorShapeProto.Dim
            for n2, _, v2 in _fields(bytes(val)):
                if n2 == 1:
                    dims.append(_signed(v2))
    return dims


def _parse_tensor(buf) -> np.ndarray:
    dtype, shape, content = 0, [], None
    floats: List[float] = []
    halfs: List[int] = []
    ints: List[int] = []
    strings: List[bytes] = []
    for num, wt, val in _fields(bytes(buf)):
        if num == 1:
            dtype = val
        elif num == 2:
            shape = _parse_shape(val)
        elif num == 4:
            content = bytes(val)
        elif num == 5:                                  # float_val (packed or not)
            floats.extend(struct.unpack(f'<{len(val) // 4}f', bytes(val)) if wt == 2 else struct.unpack('<f', val))
        elif num == 13:                                 # half_val: uint16 bit patterns as int32
            halfs.extend(_packed_varints(val) if wt == 2 else [val])
        elif num in (7, 10):                            # int_val / int64_val
            ints.extend(_packed_varints(val) if wt == 2 else [_signed(val)])
        elif num == 8:
            strings.append(bytes(val))
    n = int(np.prod(shape)) if shape else 1
    if dtype == DT_STRING:
        return np.array([s.decode('utf-8', 'replace') for s in strings], dtype=object).reshape(shape or [len(strings)])
    np_dt = {DT_FLOAT: np.float32, DT_HALF: np.float16, DT_INT32: np.int32, DT_INT64: np.int64}.get(dtype)
    if np_dt is None:
        raise ValueError(f'unsupported tensor dtype {dtype}')
    if content is not None and len(content):
        arr = np.frombuffer(content, dtype=np_dt).copy()
    elif dtype == DT_FLOAT:
        arr = np.asarray(floats, np.float32)
    elif dtype == DT_HALF:
        arr = np.asarray(halfs, np.uint16).view(np.float16)
    else:
        arr = np.asarray(ints, np_dt)
    if arr.size == 1 and n > 1:                         # a single value stands for a constant-filled tensor
        arr = np.full(n, arr[0], dtype=arr.dtype)
    return arr.reshape(shape)


class Node:
    __slots__ = ('name', 'op', 'inputs', 'attr')

    def __init__(self):
        self.name, self.op, self.inputs, self.attr = '', '', [], {}


def _parse_attr(buf):
    """AttrValue -> python value for the members the importer needs (tensor, list(i), s, i)."""
    for num, wt, val in _fields(bytes(buf)):
        if num == 8:
            return _parse_tensor(val)
        if num == 2:
            return bytes(val)
        if num == 3:
            return _signed(val)
        if num == 1:                                    # ListValue: i = field 3
            out = []
            for n2, w2, v2 in _fields(bytes(val)):
                if n2 == 3:
                    out.extend(_packed_varints(v2) if w2 == 2 else [_signed(v2)])
            return out
    return None


def parse_graph_def(data: bytes) -> 'OrderedDict[str, Node]':
    nodes: 'OrderedDict[str, Node]' = OrderedDict()
    for num, _, val in _fields(data):
        if num != 1:                                    # GraphDef.node
            continue
        node = Node()
        for n2, _, v2 in _fields(bytes(val)):
            if n2 == 1:
                node.name = bytes(v2).decode()
            elif n2 == 2:
                node.op = bytes(v2).decode()
            elif n2 == 3:
                node.inputs.append(bytes(v2).decode())
            elif n2 == 5:                               # map<string, AttrValue> entry
                key, value = None, None
                for n3, _, v3 in _fields(bytes(v2)):
                    if n3 == 1:
                        key = bytes(v3).decode()
                    elif n3 == 2:
                        value = v3
                if key in ('value', 'strides', 'dilations', 'data_format', 'padding') and value is not None:
                    node.attr[key] = _parse_attr(value)
        nodes[node.name] = node
    return nodes


# ---------------------------------------------------------------------------------------------
# graph walk
# ---------------------------------------------------------------------------------------------
def _const_of(nodes: Dict[str, Node], ref: str) -> Optional[np.ndarray]:
    """Follows Identity / Cast / ... from an input reference to the Const that feeds it."""
    for _ in range(16):
        name = ref.lstrip('^').split(':')[0]
        node = nodes.get(name)
        if node is None:
            return None
        if node.op == 'Const':
            return node.attr.get('value')
        if node.op in _PASS_THROUGH and node.inputs:
            ref = node.inputs[0]
            continue
        return None
    return None


def _const_name_of(nodes: Dict[str, Node], ref: str) -> Optional[str]:
    """Name of the Const node an input reference resolves to (through Identity / Cast ...)."""
    for _ in range(16):
        name = ref.lstrip('^').split(':')[0]
        node = nodes.get(name)
        if node is None:
            return None
        if node.op == 'Const':
            return name
        if node.op in _PASS_THROUGH and node.inputs:
            ref = node.inputs[0]
            continue
        return None
    return None


class FrozenModel:
    """What a frozen MeTRo graph contains, in this repository's terms."""

    def __init__(self):
        self.arch = ''
        self.stride = 0
        self.n_joints_model = 0
        self.depth = 8
        self.centered_stride = True
        self.permutation: List[int] = []
        self.joint_names: List[str] = []
        self.joint_edges = np.zeros((0, 2), np.int64)
        self.weights: 'OrderedDict[str, np.ndarray]' = OrderedDict()

    @property
    def spec(self) -> NetSpec:
        return NetSpec(self.arch, self.stride, self.n_joints_model, centered_stride=self.centered_stride)


def import_frozen_graph(data: bytes, depth: int = 8) -> FrozenModel:
    nodes = parse_graph_def(data)
    m = FrozenModel()
    m.depth = depth
    found: Dict[str, np.ndarray] = {}
    decomposed: Dict[str, Dict[str, np.ndarray]] = {}     # BN scope -> {'scale', 'offset'} of a non-fused batch norm
    stride_product = 1
    strided_padding: Dict[str, str] = {}                  # scope of every stride-2 3x3 convolution -> its padding attribute
    for node in nodes.values():
        if '/resnet_v2_' not in node.name:
            continue
        pre, rest = node.name.split('/resnet_v2_', 1)
        arch_num, _, local = rest.partition('/')
        arch = f'resnet_v2_{arch_num}'
        if m.arch and m.arch != arch:
            raise ValueError(f'graph mixes {m.arch} and {arch}')
        m.arch = arch
        scope = local.rsplit('/', 1)[0]
        if node.op == 'Conv2D':
            w = _const_of(nodes, node.inputs[1])
            if w is None:
                raise ValueError(f'{node.name}: filter is not a constant')
            # graph transforms may rename the convolution (fold_batch_norms gives it the name of the multiplication it
            # absorbed): when the filter constant still carries the variable's name, that name decides the scope
            wname = _const_name_of(nodes, node.inputs[1]) or ''
            if wname.endswith('/weights') and '/resnet_v2_' in wname:
                scope = wname.split('/resnet_v2_', 1)[1].partition('/')[2].rsplit('/', 1)[0]
            found[f'{scope}/weights'] = np.asarray(w, np.float32)
            s = node.attr.get('strides') or [1]
            stride_product *= max(s)
            pad = node.attr.get('padding')
            if max(s) == 2 and np.asarray(w).shape[0] == 3 and isinstance(pad, (bytes, bytearray)):
                strided_padding[scope] = bytes(pad).decode()
        elif node.op == 'BiasAdd':
            b = _const_of(nodes, node.inputs[1])
            if b is None:
                raise ValueError(f'{node.name}: bias is not a constant')
            found[f'{scope}/biases'] = np.asarray(b, np.float32)
        elif node.op.startswith('FusedBatchNorm'):
            # scope is '<conv>/BatchNorm', '<unit>/preact' or 'postnorm'
            for leaf, ref in zip(('gamma', 'beta', 'moving_mean', 'moving_variance'), node.inputs[1:5]):
                v = _const_of(nodes, ref)
                if v is None:
                    raise ValueError(f'{node.name}: {leaf} is not a constant')
                found[f'{scope}/{leaf}'] = np.asarray(v, np.float32)
        elif node.op in ('Mul', 'Add', 'AddV2') and '/batchnorm/' in local:
            # a batch norm built from primitive ops, after fold_constants (+ fold_batch_norms behind a convolution):
            # `<scope>/batchnorm/mul_1` = x * scale, `<scope>/batchnorm/add_1` = ... + offset
            bn_scope, leaf = local.split('/batchnorm/', 1)
            consts = [c for c in (_const_of(nodes, r) for r in node.inputs[:2]) if c is not None]
            if leaf == 'mul_1' and node.op == 'Mul' and len(consts) == 1:
                decomposed.setdefault(bn_scope, {})['scale'] = np.asarray(consts[0], np.float32).reshape(-1)
            elif leaf == 'add_1' and node.op != 'Mul' and len(consts) == 1:
                decomposed.setdefault(bn_scope, {})['offset'] = np.asarray(consts[0], np.float32).reshape(-1)
        elif node.op in ('MaxPool',):
            s = node.attr.get('strides') or [1]
            if 'pool1' in node.name:
                stride_product *= max(s)
    if not m.arch:
        raise ValueError('no resnet_v2_50 / resnet_v2_101 scope in the graph')
    # A decomposed batch norm y = x * scale + offset (scale already inside the filter when it followed a
    # convolution) is expressed in the blob's terms as gamma = scale, beta = offset, mean = 0 and a variance that
    # makes gamma / sqrt(variance + eps) == gamma: metro_create folds the four vectors back into scale / shift.
    for bn_scope, so in decomposed.items():
        if f'{bn_scope}/gamma' in found or 'offset' not in so:
            continue
        off = so['offset']
        found[f'{bn_scope}/gamma'] = so.get('scale', np.ones_like(off))
        found[f'{bn_scope}/beta'] = off
        found[f'{bn_scope}/moving_mean'] = np.zeros_like(off)
        found[f'{bn_scope}/moving_variance'] = np.full_like(off, 1.0 - BN_EPS)
    if 'logits/weights' not in found:
        raise ValueError('no logits convolution in the graph')
    head = found['logits/weights'].shape[3]
    if head % depth:
        raise ValueError(f'head width {head} is not a multiple of depth {depth}')
    m.n_joints_model = head // depth
    m.stride = stride_product
    spec = m.spec                                        # raises ValueError like the reference for a bad stride
    # FLAGS.centered_stride (src/options.py:118) is not stored in the graph, but it shows: conv2d_same emits a SAME
    # convolution for the strided unit of the centred block and an explicit pad + VALID convolution for every other
    # strided unit (resnet_utils.py:120-135, resnet_v2.py:277-302).  A --no-centered-stride export has VALID everywhere.
    if strided_padding:
        def pattern(sp):
            return {f'{u.name}/bottleneck_v2/conv2': ('SAME' if u.shift else 'VALID') for u in sp.units if u.stride == 2}
        if strided_padding == pattern(spec):
            pass
        elif strided_padding == pattern(NetSpec(m.arch, m.stride, m.n_joints_model, centered_stride=False)):
            m.centered_stride = False
            spec = m.spec
        else:
            raise ValueError(f'the padding of the strided convolutions ({strided_padding}) matches neither a centred-stride nor a '
                             'plain export of this architecture / stride')
    for name, shape in blob_order(spec):
        if name not in found:
            raise ValueError(f'the graph has no constant for {name}')
        a = found[name]
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f'{name}: graph has shape {a.shape}, the architecture needs {shape}')
        m.weights[name] = a
    # fetches (src/main.py:127,140-141)
    out = nodes.get('output')
    if out is not None and len(out.inputs) >= 2:
        idx = _const_of(nodes, out.inputs[1])
        if idx is not None:
            m.permutation = [int(v) for v in np.asarray(idx).reshape(-1)]
    if not m.permutation:
        m.permutation = list(range(m.n_joints_model))
    jn = nodes.get('joint_names')
    if jn is not None and jn.op == 'Const':
        m.joint_names = [str(s) for s in np.asarray(jn.attr['value']).reshape(-1)]
    je = nodes.get('joint_edges')
    if je is not None and je.op == 'Const':
        m.joint_edges = np.asarray(je.attr['value'], np.int64).reshape(-1, 2)
    return m


def load_frozen_model(path: str, depth: int = 8) -> FrozenModel:
    with open(path, 'rb') as f:
        return import_frozen_graph(f.read(), depth)
