"""A numpy/torch-CPU stand-in for the few TensorFlow 1.13 / tf.contrib.slim entry points the
reference's inference graph builders call.  TEST INFRASTRUCTURE (oracle/): it exists so that the
REFERENCE'S OWN PYTHON FILES can be imported from /root/reference and executed here, eagerly, in
float64 -- TensorFlow itself cannot be installed in this container (no wheel, no network, no
Python 3.12 build of TF 1.13).

What this pins and what it does not
    * Executed verbatim from the reference: src/model/architectures.py, src/model/resnet_v2.py,
      src/model/resnet_utils.py, src/model/volumetric.py (build_inference_model,
      net_output_to_heatmap_and_coords, heatmap_to_metric), src/tfu.py (softmax, decode_heatmap,
      layout converters, static_* helpers), src/tfu3d.py (root_relative), src/data/datasets.py
      (JointInfo.permute_joints), src/util.py (invert_permutation), src/options.py (flag defaults).
      So block tables, centred-stride selection, stride/atrous bookkeeping, padding decisions,
      shortcut wiring, variable names/shapes/creation order, the reshape/transpose/softmax/decode
      chain, metric scaling and root subtraction all come from the reference's code.
    * Restated here from TensorFlow's documented semantics (the TF op kernels are a third-party
      dependency, tensorflow-gpu==1.13.1, install_dependencies.sh:13): conv2d with SAME/VALID
      padding, stride and dilation (HWIO filters); FusedBatchNorm in inference form; max_pool2d
      with SAME/VALID padding; Pad; elementwise/reduction ops.  Each is a few lines below.

Only oracle/gen_golden.py (and tests that exercise it when /root/reference is present) use this.
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# dtypes / shapes / tensors
# ------------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, np_dtype):
        self.name, self.as_numpy_dtype = name, np_dtype
        self.is_floating = name.startswith('float')

    def __repr__(self):
        return f'tf.{self.name}'


float16, float32, float64 = DType('float16', np.float16), DType('float32', np.float32), DType('float64', np.float64)
int32, int64, bool_, string = DType('int32', np.int32), DType('int64', np.int64), DType('bool', np.bool_), DType('string', np.str_)


def as_dtype(d):
    if isinstance(d, DType):
        return d
    d = np.dtype(d)
    for t in (float16, float32, float64, int32, int64, bool_):
        if np.dtype(t.as_numpy_dtype) == d:
            return t
    return string


class TensorShape:
    def __init__(self, dims):
        self.dims = list(dims)
        self.ndims = len(self.dims)

    def as_list(self):
        return list(self.dims)

    def __len__(self):
        return self.ndims

    def __getitem__(self, i):
        return self.dims[i]

    def __iter__(self):
        return iter(self.dims)


class Tensor:
    """Eager tensor: float data is carried in float64 whatever its nominal dtype (the golden vectors
    are the exact mathematical function of the graph); `dtype` records the nominal TF dtype."""

    def __init__(self, value, dtype: Optional[DType] = None, name: Optional[str] = None):
        value = np.asarray(value)
        if dtype is None:
            dtype = as_dtype(value.dtype)
        if dtype.is_floating:
            value = value.astype(np.float64)
        self.value, self.dtype, self.name = value, dtype, name

    def get_shape(self):
        return TensorShape(self.value.shape)

    @property
    def shape(self):
        return TensorShape(self.value.shape)

    def set_shape(self, s):
        pass

    def _bin(self, other, fn, rev=False):
        o = other.value if isinstance(other, Tensor) else np.asarray(other, dtype=np.float64)
        return Tensor(fn(o, self.value) if rev else fn(self.value, o), self.dtype)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __neg__(self): return Tensor(-self.value, self.dtype)
    def __gt__(self, o): return Tensor(np.greater(self.value, _v(o)), bool_)

    def __getitem__(self, idx):
        if isinstance(idx, list):
            idx = tuple(idx)
        return Tensor(self.value[idx], self.dtype)


def _v(x):
    return x.value if isinstance(x, Tensor) else np.asarray(x)


def convert_to_tensor(x, dtype=None, name=None):
    if isinstance(x, Tensor):
        return Tensor(x.value, dtype or x.dtype, name)
    return Tensor(np.asarray(x), dtype, name)


def _axes(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else int(axis)


# ------------------------------------------------------------------------------------------------
# variables and scopes
# ------------------------------------------------------------------------------------------------
class _State:
    def __init__(self):
        self.scopes: List[str] = []
        self.variables: Dict[str, np.ndarray] = {}
        self.prefix = ''
        self.created: List[tuple] = []       # (full name, shape) in creation order
        self.arg_scopes: List[dict] = [{}]
        self.trace: Optional[Dict[str, np.ndarray]] = None


_S = _State()


def reset(variables: Dict[str, np.ndarray], strip_prefix: str = ''):
    """Install the weight dictionary the graph builders will read through get_variable."""
    _S.scopes, _S.variables, _S.prefix, _S.created = [], dict(variables), strip_prefix, []
    _S.arg_scopes = [{}]
    _S.trace = None


def created_variables():
    return list(_S.created)


def get_variable(name, shape):
    full = '/'.join(_S.scopes + [name])
    key = full[len(_S.prefix):] if full.startswith(_S.prefix) else full
    if key not in _S.variables:
        raise KeyError(f'graph asked for variable {full!r} (key {key!r}) which the weight set does not hold')
    a = np.asarray(_S.variables[key])
    if tuple(a.shape) != tuple(shape):
        raise ValueError(f'variable {full}: graph wants shape {tuple(shape)}, weight set has {a.shape}')
    _S.created.append((key, tuple(shape)))
    return a.astype(np.float64)


class _Scope:
    def __init__(self, name):
        self.name = name
        self.original_name_scope = name + '/'


@contextlib.contextmanager
def variable_scope(scope=None, default_name=None, values=None, reuse=None, custom_getter=None, **kw):
    name = scope if scope is not None else default_name
    if isinstance(name, _Scope):
        name = name.name.split('/')[-1]
    _S.scopes.append(name)
    try:
        yield _Scope('/'.join(_S.scopes))
    finally:
        _S.scopes.pop()


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name or default_name


# ------------------------------------------------------------------------------------------------
# arg_scope (tf.contrib.framework): defaults for decorated ops, nested scopes override outer ones
# ------------------------------------------------------------------------------------------------
def add_arg_scope(func):
    key = (func.__module__, func.__qualname__)

    def wrapper(*args, **kwargs):
        merged = dict(_S.arg_scopes[-1].get(key, {}))
        merged.update(kwargs)
        return func(*args, **merged)

    wrapper._arg_scope_key = key
    wrapper.__name__ = func.__name__
    wrapper.__wrapped__ = func
    return wrapper


@contextlib.contextmanager
def arg_scope(list_ops_or_scope, **kwargs):
    if isinstance(list_ops_or_scope, dict):
        if kwargs:
            raise ValueError('When attempting to re-use a scope by suppling a dictionary, kwargs must be empty.')
        # re-using a captured scope replaces the current defaults (tf.contrib.framework.arg_scope)
        new = {k: dict(v) for k, v in list_ops_or_scope.items()}
    else:
        new = {k: dict(v) for k, v in _S.arg_scopes[-1].items()}
        for op in list_ops_or_scope:
            key = op._arg_scope_key
            cur = dict(new.get(key, {}))
            cur.update(kwargs)
            new[key] = cur
    _S.arg_scopes.append(new)
    try:
        yield new
    finally:
        _S.arg_scopes.pop()


# ------------------------------------------------------------------------------------------------
# ops (documented TensorFlow semantics)
# ------------------------------------------------------------------------------------------------
def cast(x, dtype, name=None):
    return Tensor(_v(x), dtype)


def identity(x, name=None):
    return Tensor(_v(x), x.dtype if isinstance(x, Tensor) else None, name)


def reshape(x, shape, name=None):
    return Tensor(_v(x).reshape([int(s) for s in shape]), x.dtype)


def transpose(x, perm=None, name=None):
    return Tensor(np.transpose(_v(x), perm), x.dtype)


def reduce_max(x, axis=None, keepdims=False, name=None):
    return Tensor(np.max(_v(x), axis=_axes(axis), keepdims=keepdims), x.dtype)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    return Tensor(np.sum(_v(x), axis=_axes(axis), keepdims=keepdims), x.dtype)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    return Tensor(np.mean(_v(x), axis=_axes(axis), keepdims=keepdims), x.dtype)


def exp(x, name=None):
    return Tensor(np.exp(_v(x)), x.dtype)


def linspace(start, stop, num, name=None):
    # tf.linspace includes both end points: start + i * (stop - start) / (num - 1)
    return Tensor(np.linspace(float(start), float(stop), int(num)), float32)


def squeeze(x, axis=None, name=None):
    return Tensor(np.squeeze(_v(x), axis=_axes(axis)), x.dtype)


def stack(values, axis=0, name=None):
    return Tensor(np.stack([_v(v) for v in values], axis=axis), values[0].dtype)


def concat(values, axis, name=None):
    return Tensor(np.concatenate([_v(v) for v in values], axis=axis), values[0].dtype)


def expand_dims(x, axis, name=None):
    return Tensor(np.expand_dims(_v(x), axis), x.dtype)


def gather(params, indices, axis=0, name=None):
    return Tensor(np.take(_v(params), np.asarray(indices), axis=axis), params.dtype, name)


def einsum(equation, *inputs, name=None):
    return Tensor(np.einsum(equation, *[_v(x) for x in inputs]))


def where(condition, x=None, y=None, name=None):
    """TF 1.x semantics for a rank-1 condition: it selects whole rows (first axis) of x / y."""
    c = np.asarray(_v(condition))
    xv, yv = _v(x), _v(y)
    if c.ndim == 1 and xv.ndim > 1:
        c = c.reshape((-1,) + (1,) * (xv.ndim - 1))
    return Tensor(np.where(c, xv, yv))


def det(x, name=None):
    return Tensor(np.linalg.det(_v(x)))


def ones_like(x, dtype=None, name=None):
    return Tensor(np.ones_like(_v(x)), dtype or x.dtype)


def zeros_like(x, dtype=None, name=None):
    return Tensor(np.zeros_like(_v(x)), dtype or x.dtype)


def pad(x, paddings, mode='CONSTANT', name=None, constant_values=0):
    return Tensor(np.pad(_v(x), [tuple(p) for p in paddings], mode='constant', constant_values=constant_values), x.dtype)


def relu(x, name=None):
    return Tensor(np.maximum(_v(x), 0.0), x.dtype)


def _same_pad(n, k_eff, s):
    """TensorFlow 'SAME': out = ceil(n / s); pad_total = max((out-1)*s + k_eff - n, 0); the smaller
    half goes in front."""
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def _to_nchw(x, data_format):
    return x if data_format == 'NCHW' else np.transpose(x, (0, 3, 1, 2))


def _from_nchw(x, data_format):
    return x if data_format == 'NCHW' else np.transpose(x, (0, 2, 3, 1))


def _trace(t: Tensor, data_format):
    if _S.trace is not None:
        _S.trace['/'.join(_S.scopes)] = _from_nchw(_to_nchw(t.value, data_format), 'NHWC').copy()
    return t


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else tuple(int(a) for a in v)


@add_arg_scope
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding='SAME', data_format=None, rate=1,
           activation_fn=relu, normalizer_fn=None, normalizer_params=None, weights_initializer=None,
           weights_regularizer=None, biases_initializer='zeros', biases_regularizer=None, reuse=None,
           variables_collections=None, outputs_collections=None, trainable=True, scope=None):
    """tf.contrib.layers.conv2d: variables 'weights' (HWIO) and -- iff normalizer_fn is None and
    biases_initializer is not None -- 'biases'; convolution, then normalizer or bias, then activation."""
    data_format = data_format or 'NHWC'
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)
    rh, rw = _pair(rate)
    with variable_scope(scope, 'Conv', [inputs]):
        x = _to_nchw(_v(inputs), data_format)
        cin = x.shape[1]
        w = get_variable('weights', (kh, kw, cin, num_outputs))
        if padding == 'SAME':
            ph = _same_pad(x.shape[2], kh + (kh - 1) * (rh - 1), sh)
            pw = _same_pad(x.shape[3], kw + (kw - 1) * (rw - 1), sw)
            x = np.pad(x, ((0, 0), (0, 0), ph, pw))
        elif padding != 'VALID':
            raise ValueError(padding)
        y = F.conv2d(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1))),
                     None, stride=(sh, sw), dilation=(rh, rw)).numpy()
        out = Tensor(_from_nchw(y, data_format), inputs.dtype)
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            b = get_variable('biases', (num_outputs,))
            shape = (1, -1, 1, 1) if data_format == 'NCHW' else (1, 1, 1, -1)
            out = Tensor(out.value + b.reshape(shape), out.dtype)
        if activation_fn is not None:
            out = activation_fn(out)
        return _trace(out, data_format)


def conv3d(*a, **k):
    raise NotImplementedError


conv3d = add_arg_scope(conv3d)


@add_arg_scope
def batch_norm(inputs, decay=0.999, center=True, scale=False, epsilon=0.001, activation_fn=None,
               param_initializers=None, param_regularizers=None, updates_collections=None, is_training=True,
               reuse=None, variables_collections=None, outputs_collections=None, trainable=True,
               batch_weights=None, fused=None, data_format='NHWC', zero_debias_moving_mean=False, scope=None,
               renorm=False, renorm_clipping=None, renorm_decay=0.99, adjustment=None):
    """tf.contrib.layers.batch_norm, inference form (FusedBatchNorm with is_training=False):
    y = gamma * (x - moving_mean) / sqrt(moving_variance + epsilon) + beta."""
    if is_training:
        raise NotImplementedError('the shim only evaluates inference-mode batch norm')
    data_format = data_format or 'NHWC'
    with variable_scope(scope, 'BatchNorm', [inputs]):
        x = _v(inputs)
        c = x.shape[1] if data_format == 'NCHW' else x.shape[-1]
        shape = (1, -1, 1, 1) if data_format == 'NCHW' else (1, 1, 1, -1)
        beta = get_variable('beta', (c,)) if center else np.zeros(c)
        gamma = get_variable('gamma', (c,)) if scale else np.ones(c)
        mean = get_variable('moving_mean', (c,))
        var = get_variable('moving_variance', (c,))
        y = (x - mean.reshape(shape)) / np.sqrt(var.reshape(shape) + epsilon) * gamma.reshape(shape) + beta.reshape(shape)
        out = Tensor(y, inputs.dtype)
        if activation_fn is not None:
            out = activation_fn(out)
        return _trace(out, data_format)


@add_arg_scope
def max_pool2d(inputs, kernel_size, stride=2, padding='VALID', data_format='NHWC', outputs_collections=None,
               scope=None):
    """tf.contrib.layers.max_pool2d: VALID = no padding; SAME pads (TensorFlow SAME amounts) with
    values that never win the max."""
    data_format = data_format or 'NHWC'
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)
    x = _to_nchw(_v(inputs), data_format)
    if padding == 'SAME':
        ph, pw = _same_pad(x.shape[2], kh, sh), _same_pad(x.shape[3], kw, sw)
        x = np.pad(x, ((0, 0), (0, 0), ph, pw), constant_values=-np.inf)
    y = F.max_pool2d(torch.from_numpy(np.ascontiguousarray(x)), (kh, kw), (sh, sw)).numpy()
    with variable_scope(scope, 'MaxPool2D', [inputs]):
        return _trace(Tensor(_from_nchw(y, data_format), inputs.dtype), data_format)


def _unused_layer(*a, **k):
    raise NotImplementedError('not on the inference path')


def softmax(logits, scope=None):
    e = np.exp(_v(logits) - np.max(_v(logits), axis=-1, keepdims=True))
    return Tensor(e / e.sum(axis=-1, keepdims=True), logits.dtype)


def collect_named_outputs(collections, alias, outputs):
    return outputs


def convert_collection_to_dict(collection, clear_collection=False):
    return {}


# ------------------------------------------------------------------------------------------------
# module tree
# ------------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Registers fake `tensorflow...` modules (and the two tiny third-party helpers the reference
    imports, attrdict and more_itertools) in sys.modules.  Idempotent."""
    if 'tensorflow' in sys.modules and getattr(sys.modules['tensorflow'], '_metro_shim', False):
        return sys.modules['tensorflow']
    this = sys.modules[__name__]
    nn = _module('tensorflow.nn', relu=relu)
    layer_fns = dict(
        conv2d=conv2d, conv3d=conv3d, batch_norm=batch_norm, max_pool2d=max_pool2d, softmax=softmax,
        l2_regularizer=lambda scale, scope=None: None,
        variance_scaling_initializer=lambda *a, **k: 'variance_scaling',
        arg_scope=arg_scope, add_arg_scope=add_arg_scope)
    for n in ('conv3d_transpose', 'conv2d_transpose', 'avg_pool2d', 'separable_conv2d', 'spatial_softmax'):
        f = types.FunctionType(_unused_layer.__code__, globals(), n)
        f.__qualname__ = n
        layer_fns[n] = add_arg_scope(f)
    utils = _module('tensorflow.contrib.layers.python.layers.utils', collect_named_outputs=collect_named_outputs,
                    convert_collection_to_dict=convert_collection_to_dict)
    layers_mod = _module('tensorflow.contrib.layers.python.layers.layers', **layer_fns)
    initializers = _module('tensorflow.contrib.layers.python.layers.initializers',
                           variance_scaling_initializer=layer_fns['variance_scaling_initializer'])
    regularizers = _module('tensorflow.contrib.layers.python.layers.regularizers', l2_regularizer=layer_fns['l2_regularizer'])
    pl = _module('tensorflow.contrib.layers.python.layers', layers=layers_mod, utils=utils, initializers=initializers,
                 regularizers=regularizers)
    lp = _module('tensorflow.contrib.layers.python', layers=pl)
    contrib_layers = _module('tensorflow.contrib.layers', python=lp, **layer_fns)
    slim = _module('tensorflow.contrib.slim', **layer_fns)
    fpo = _module('tensorflow.contrib.framework.python.ops', add_arg_scope=add_arg_scope, arg_scope=arg_scope)
    fp = _module('tensorflow.contrib.framework.python', ops=fpo)
    framework = _module('tensorflow.contrib.framework', python=fp, add_arg_scope=add_arg_scope, arg_scope=arg_scope)
    contrib = _module('tensorflow.contrib', slim=slim, layers=contrib_layers, framework=framework)
    math_ops = _module('tensorflow.python.ops.math_ops', reduce_mean=reduce_mean)
    nn_ops = _module('tensorflow.python.ops.nn_ops', relu=relu)
    vs = _module('tensorflow.python.ops.variable_scope', variable_scope=variable_scope)
    array_ops = _module('tensorflow.python.ops.array_ops', pad=pad)
    pops = _module('tensorflow.python.ops', math_ops=math_ops, nn_ops=nn_ops, variable_scope=vs, array_ops=array_ops)
    fops = _module('tensorflow.python.framework.ops')
    pframework = _module('tensorflow.python.framework', ops=fops)
    python = _module('tensorflow.python', ops=pops, framework=pframework)
    tf = _module(
        'tensorflow', _metro_shim=True, __version__='1.13.1-metro-shim', Tensor=Tensor, TensorShape=TensorShape,
        float16=float16, float32=float32, float64=float64, int32=int32, int64=int64, bool=bool_, string=string,
        as_dtype=as_dtype, convert_to_tensor=convert_to_tensor, cast=cast, identity=identity, reshape=reshape,
        transpose=transpose, reduce_max=reduce_max, reduce_sum=reduce_sum, reduce_mean=reduce_mean, exp=exp,
        linspace=linspace, squeeze=squeeze, stack=stack, concat=concat, expand_dims=expand_dims, gather=gather,
        ones_like=ones_like, zeros_like=zeros_like, pad=pad, einsum=einsum, where=where,
        linalg=_module('tensorflow.linalg', det=det), variable_scope=variable_scope, name_scope=name_scope,
        nn=nn, contrib=contrib, python=python, shim=this)

    class AttrDict(dict):
        """attrdict.AttrDict stand-in (third-party helper): attribute access to dict items."""
        __getattr__ = dict.__getitem__
        __setattr__ = dict.__setitem__

    _module('attrdict', AttrDict=AttrDict)

    def pairwise(it):
        it = list(it)
        return zip(it[:-1], it[1:])

    _module('more_itertools', pairwise=pairwise)
    return tf
