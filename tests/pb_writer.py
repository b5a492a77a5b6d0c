"""Minimal encoder of the GraphDef messages the importer reads (TEST INFRASTRUCTURE): builds a frozen graph
with the node names, ops and attributes the reference's export produces (src/main.py:106-160,
src/model/resnet_v2.py) from a weight dictionary, without TensorFlow."""
import struct

import numpy as np


def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num: int, payload: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _vi(num: int, v: int) -> bytes:
    return _varint(num << 3) + _varint(v)


def tensor(arr, form='content') -> bytes:
    arr = np.asarray(arr)
    shape = b''.join(_ld(2, _vi(1, d)) for d in arr.shape)
    if arr.dtype == object or arr.dtype.kind in 'US':
        body = _vi(1, 7) + _ld(2, shape) + b''.join(_ld(8, str(s).encode()) for s in arr.reshape(-1))
        return body
    dt = {np.dtype(np.float32): 1, np.dtype(np.float16): 19, np.dtype(np.int32): 3, np.dtype(np.int64): 9}[arr.dtype]
    body = _vi(1, dt) + _ld(2, shape)
    if form == 'content':
        body += _ld(4, np.ascontiguousarray(arr).tobytes())
    elif dt == 1:
        body += _ld(5, struct.pack(f'<{arr.size}f', *arr.reshape(-1)))
    elif dt == 19:
        body += _ld(13, b''.join(_varint(int(v)) for v in arr.reshape(-1).view(np.uint16)))
    else:
        body += _ld(7 if dt == 3 else 10, b''.join(_varint(int(v)) for v in arr.reshape(-1)))
    return body


def attr_tensor(arr, form='content') -> bytes:
    return _ld(8, tensor(arr, form))


def attr_ints(vals) -> bytes:
    return _ld(1, _ld(3, b''.join(_varint(v) for v in vals)))


def attr_str(s: str) -> bytes:
    return _ld(2, s.encode())


def node(name, op, inputs=(), **attrs) -> bytes:
    body = _ld(1, name.encode()) + _ld(2, op.encode())
    for i in inputs:
        body += _ld(3, i.encode())
    for k, v in attrs.items():
        body += _ld(5, _ld(1, k.encode()) + _ld(2, v))
    return _ld(1, body)


def frozen_graph(spec, weights, permutation, joint_names, joint_edges, half=False, prefix='MainPart', bn_form='fused',
                 eps=1e-5, fold_renames=False):
    """GraphDef bytes.  Constants get fold_constants-style names (not the variable names) and reach their
    consumers through Identity nodes; fp16 graphs store half constants, alternating the two TensorProto
    encodings (tensor_content / typed value list).

    bn_form='fused'  : FusedBatchNorm nodes with their four constants (what slim's fused batch norm exports).
    bn_form='folded' : what fold_constants + fold_batch_norms leave of a batch norm built from primitive ops
                       (tf.nn.batch_normalization): after a convolution the scale is folded into the filter and
                       only `<scope>/batchnorm/add_1` (Add with a constant) remains; a batch norm that does not
                       follow a convolution (preact, postnorm) is `<scope>/batchnorm/mul_1` + `.../add_1`."""
    out = [node('input', 'Placeholder')]
    root = f'{prefix}/{spec.arch}'
    counter = [0]

    def const(value, name=None):
        counter[0] += 1
        name = name or f'{root}/_cf_{counter[0]}'
        v = np.asarray(value, np.float16 if half else np.float32)
        form = 'content' if (counter[0] % 2 or v.size > 4096) else 'list'
        out.append(node(name, 'Const', value=attr_tensor(v, form)))
        out.append(node(name + '/read', 'Identity', [name]))
        return name + '/read'

    def scale_offset(scope):
        g, b, m, v = (np.asarray(weights[f'{scope}/{leaf}'], np.float64)
                      for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'))
        sc = g / np.sqrt(v + eps)
        return sc, b - m * sc

    def conv(c, scope, src):
        s = c.stride
        w = weights[f'{scope}/weights']
        fold = bn_form == 'folded' and c.has_bn
        if fold:
            sc, off = scale_offset(f'{scope}/BatchNorm')
            w = np.asarray(w, np.float64) * sc                       # HWIO: the scale runs along the output channels
        # conv2d_same (resnet_utils.py:120-135): SAME for stride 1 and for centred strides, explicit pad + VALID otherwise
        same = c.stride == 1 or (c.k > 1 and c.pad_lo + c.pad_hi < c.k + (c.k - 1) * (c.rate - 1) - 1) or c.k == 1
        # fold_renames: the rewritten convolution carries the name of the multiplication it absorbed, while its filter
        # constant keeps the variable's name -- the naming a graph-transform tool may leave behind
        cname = f'{root}/{scope}/BatchNorm/batchnorm/mul_1' if (fold and fold_renames) else f'{root}/{scope}/Conv2D'
        out.append(node(cname, 'Conv2D', [src, const(w, f'{root}/{scope}/weights' if fold_renames else None)],
                        strides=attr_ints([1, 1, s, s]), data_format=attr_str('NCHW'),
                        padding=attr_str('SAME' if same else 'VALID')))
        last = cname
        if c.has_bias:
            out.append(node(f'{root}/{scope}/BiasAdd', 'BiasAdd', [last, const(weights[f'{scope}/biases'])]))
            last = f'{root}/{scope}/BiasAdd'
        if fold:
            out.append(node(f'{root}/{scope}/BatchNorm/batchnorm/add_1', 'Add', [last, const(off)]))
            last = f'{root}/{scope}/BatchNorm/batchnorm/add_1'
        elif c.has_bn:
            last = bn(f'{scope}/BatchNorm', last)
        return last

    def bn(scope, src):
        if bn_form == 'folded':
            sc, off = scale_offset(scope)
            out.append(node(f'{root}/{scope}/batchnorm/mul_1', 'Mul', [src, const(sc)]))
            out.append(node(f'{root}/{scope}/batchnorm/add_1', 'Add', [f'{root}/{scope}/batchnorm/mul_1', const(off)]))
            return f'{root}/{scope}/batchnorm/add_1'
        refs = [const(weights[f'{scope}/{leaf}']) for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance')]
        out.append(node(f'{root}/{scope}/FusedBatchNorm', 'FusedBatchNorm', [src] + refs))
        return f'{root}/{scope}/FusedBatchNorm'

    x = conv(spec.root, 'conv1', 'input')
    out.append(node(f'{root}/pool1/MaxPool', 'MaxPool', [x], strides=attr_ints([1, 1, 2, 2])))
    x = f'{root}/pool1/MaxPool'
    for u in spec.units:
        s = f'{u.name}/bottleneck_v2'
        pre = bn(f'{s}/preact', x)
        if u.shortcut is not None:
            sc = conv(u.shortcut, f'{s}/shortcut', pre)
        else:
            out.append(node(f'{root}/{s}/shortcut/MaxPool', 'MaxPool', [x], strides=attr_ints([1, 1, u.stride, u.stride])))
            sc = f'{root}/{s}/shortcut/MaxPool'
        r = conv(u.conv1, f'{s}/conv1', pre)
        r = conv(u.conv2, f'{s}/conv2', r)
        r = conv(u.conv3, f'{s}/conv3', r)
        out.append(node(f'{root}/{s}/add', 'Add', [sc, r]))
        x = f'{root}/{s}/add'
    x = bn('postnorm', x)
    x = conv(spec.logits, 'logits', x)
    out.append(node('perm', 'Const', value=attr_tensor(np.asarray(permutation, np.int32))))
    out.append(node('axis', 'Const', value=attr_tensor(np.asarray(1, np.int32))))
    out.append(node('output', 'GatherV2', [x, 'perm', 'axis']))
    out.append(node('joint_names', 'Const', value=attr_tensor(np.asarray(joint_names, dtype=object))))
    out.append(node('joint_edges', 'Const', value=attr_tensor(np.asarray(joint_edges, np.int64))))
    return b''.join(out)
