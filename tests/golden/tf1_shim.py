"""A small EAGER stand-in for the TensorFlow 1.x API surface that the reference NAR module uses
(/root/reference/nar_module/nar: nar_model.py, datasets.py, nar_trainer_gcom.py's model_fn), so that the reference's OWN
code can be imported and executed in this container (TensorFlow 1.12 cannot be installed: python 3.12, no network).
Users: tests/golden/make_model_golden.py (whole model graph), make_sampler_tf_golden.py (sampler distribution),
make_hook_golden.py (SessionRunHook), make_dataset_golden.py (tf.data input pipeline), make_model_fn_golden.py (model_fn).

What this is for: tests/golden/make_model_golden.py installs this module as ``tensorflow``, imports the reference's
``NARModuleModel`` unmodified and runs its constructor; every ``tf.*`` call computes immediately on torch-CPU tensors
(float64 when ``configure(float64=True)``), so the constructor's "graph" is the actual forward pass, and
``AdamOptimizer.compute_gradients`` differentiates it with torch autograd.  The results (negatives, intermediates the
reference sends to ``tf.summary.histogram``, logits, loss, gradients) become golden vectors for oracle/nar_oracle.py.

What it pins and what it does not: the WIRING of the model is the reference's (which features are concatenated in which
order, masks, normalisation statistics, scopes / variable sharing, which layers are regularised, loss normalisation...).
The SEMANTICS of each individual op below are this file's restatement of the TensorFlow 1.12 documentation (e.g. tf.unique
keeps first-occurrence order, tf.nn.moments is the biased variance, tf.nn.leaky_relu alpha 0.2, UGRNNCell /
dynamic_rnn / AdamOptimizer formulas) - they are library behaviour, not reference code.

Test infrastructure only (like oracle/): nothing in the product imports it.
"""
from __future__ import annotations

import collections
import contextlib
import sys
import types

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------- global state


class _State:
    def __init__(self):
        self.reset()
        self.float = torch.float32

    def reset(self):
        self.scope = []                      # variable-scope name stack
        self.scope_init = [None]             # default initializer stack
        self.vars = collections.OrderedDict()   # full name -> leaf tensor (requires_grad)
        self.regs = collections.OrderedDict()   # full name -> regulariser callable
        self.preset = {}                     # full name -> numpy value used instead of the initializer
        self.feeds = {}                      # placeholder name -> value or list of values (consumed in creation order)
        self.rng = torch.Generator().manual_seed(0)          # tf.random_shuffle / dropout
        self.rng_init = torch.Generator().manual_seed(1)     # variable initializers
        self.hist = []                       # (name, tensor) from tf.summary.histogram
        self.scalars = []                    # (name, tensor) from tf.summary.scalar
        self.softmax_inputs = []             # inputs of tf.nn.softmax in call order
        self.dropout_masks = []              # keep-masks of tf.layers.dropout / DropoutWrapper in call order
        self.grads = None                    # name -> gradient (AdamOptimizer.compute_gradients)
        self.vars_after = None               # name -> value after apply_gradients
        self.adam = None


S = _State()


def configure(float64=True, seed=0, feeds=None, preset=None):
    S.reset()
    S.float = torch.float64 if float64 else torch.float32
    S.rng = torch.Generator().manual_seed(seed)
    S.rng_init = torch.Generator().manual_seed(seed + 1000)
    S.feeds = dict(feeds or {})
    S.preset = dict(preset or {})
    _refresh_dtypes()


def _scope_name():
    return '/'.join(S.scope)


# ---------------------------------------------------------------------------------------------- helpers
def _int(x):
    if isinstance(x, torch.Tensor):
        return int(x.item())
    return int(x)


def _shape(shape):
    if isinstance(shape, torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [_int(v) for v in shape]


def _t(x, dtype=None):
    """python / numpy / tensor -> tensor (python floats become the configured float type, ints int32 like tf.constant)"""
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        if t.dtype in (torch.float32, torch.float64):
            t = t.to(S.float)
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)) and any(isinstance(v, torch.Tensor) for v in x):
        return torch.stack([_t(v) for v in x]) if dtype is None else torch.stack([_t(v) for v in x]).to(dtype)
    if dtype is None:
        probe = np.asarray(x)
        if probe.dtype.kind == 'f':
            dtype = S.float
        elif probe.dtype.kind == 'b':
            dtype = torch.bool
        else:
            dtype = torch.int32
    return torch.as_tensor(np.asarray(x), dtype=dtype)


class TensorShape(list):
    def as_list(self):
        return list(self)


def _get_shape(self):
    return TensorShape(int(d) for d in self.shape)


torch.Tensor.get_shape = _get_shape
torch.Tensor.set_shape = lambda self, shape: None


# ---------------------------------------------------------------------------------------------- module skeleton
def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    return m


tf = _mod('tensorflow')
tf.contrib = _mod('tensorflow.contrib')
tf.contrib.layers = _mod('tensorflow.contrib.layers')
tf.contrib.rnn = _mod('tensorflow.contrib.rnn')
tf.contrib.metrics = _mod('tensorflow.contrib.metrics')
tf.contrib.lookup = _mod('tensorflow.contrib.lookup')
tf.python = _mod('tensorflow.python')
tf.python.ops = _mod('tensorflow.python.ops')
for _n in ('control_flow_ops', 'array_ops', 'math_ops'):
    setattr(tf.python.ops, _n, _mod('tensorflow.python.ops.' + _n))
tf.nn = _mod('tensorflow.nn')
tf.nn.rnn_cell = _mod('tensorflow.nn.rnn_cell')
tf.layers = _mod('tensorflow.layers')
tf.summary = _mod('tensorflow.summary')
tf.logging = _mod('tensorflow.logging')
tf.estimator = _mod('tensorflow.estimator')
tf.train = _mod('tensorflow.train')
tf.losses = _mod('tensorflow.losses')
tf.metrics = _mod('tensorflow.metrics')
tf.initializers = _mod('tensorflow.initializers')
tf.random = _mod('tensorflow.random')

tf.int32, tf.int64, tf.bool, tf.string = torch.int32, torch.int64, torch.bool, 'string'
tf.AUTO_REUSE = 'AUTO_REUSE'
tf.TensorShape = TensorShape


def _refresh_dtypes():
    tf.float32 = S.float          # the whole reference graph runs in S.float (float64 for tight golden vectors)
    tf.float64 = torch.float64


_refresh_dtypes()


class _ModeKeys:
    TRAIN, EVAL, PREDICT = 'train', 'eval', 'infer'


tf.estimator.ModeKeys = _ModeKeys


class _GraphKeys:
    UPDATE_OPS = 'update_ops'


tf.GraphKeys = _GraphKeys
tf.get_collection = lambda *a, **k: []
tf.control_dependencies = lambda *a, **k: contextlib.nullcontext()
tf.device = lambda *a, **k: contextlib.nullcontext()
for _n in ('info', 'warn', 'warning', 'error', 'debug'):
    setattr(tf.logging, _n, lambda *a, **k: None)
tf.logging.INFO = 20
tf.logging.set_verbosity = lambda *a, **k: None


# ---------------------------------------------------------------------------------------------- summaries (recorded)
def _histogram(name, values=None, family=None, **k):
    if values is not None:
        S.hist.append((name, values.detach().clone()))


def _scalar(name, tensor=None, family=None, **k):
    if tensor is not None:
        S.scalars.append((name, _t(tensor).detach().clone()))


tf.summary.histogram = _histogram
tf.summary.scalar = _scalar

# ---------------------------------------------------------------------------------------------- scopes / variables


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, initializer=None, **k):
    name = name_or_scope if isinstance(name_or_scope, str) else (name_or_scope.name if name_or_scope is not None else default_name)
    absolute = not isinstance(name_or_scope, str) and name_or_scope is not None
    saved = list(S.scope)
    if absolute:
        S.scope[:] = name.split('/') if name else []
    elif name:
        S.scope.append(name)
    S.scope_init.append(initializer if initializer is not None else S.scope_init[-1])
    try:
        yield types.SimpleNamespace(name=_scope_name())
    finally:
        S.scope[:] = saved
        S.scope_init.pop()


tf.variable_scope = variable_scope
tf.get_variable_scope = lambda: types.SimpleNamespace(name=_scope_name())
tf.name_scope = lambda *a, **k: contextlib.nullcontext()


def _fans(shape):
    if len(shape) == 1:
        return shape[0], shape[0]
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    return shape[-2] * rf, shape[-1] * rf


def xavier_initializer(uniform=True, seed=None, dtype=None):
    def init(shape):
        fi, fo = _fans(shape)
        lim = np.sqrt(6.0 / (fi + fo))
        return (torch.rand(shape, generator=S.rng_init, dtype=torch.float64) * 2 - 1) * lim
    return init


def variance_scaling_initializer(factor=2.0, mode='FAN_IN', uniform=False, seed=None, dtype=None):
    def init(shape):
        fi, fo = _fans(shape)
        n = {'FAN_IN': fi, 'FAN_OUT': fo, 'FAN_AVG': (fi + fo) / 2.0}[mode]
        std = np.sqrt(1.3 * factor / n)
        return torch.randn(shape, generator=S.rng_init, dtype=torch.float64).clamp_(-2, 2) * std
    return init


def lecun_uniform(seed=None):
    def init(shape):
        fi, _ = _fans(shape)
        lim = np.sqrt(3.0 / fi)
        return (torch.rand(shape, generator=S.rng_init, dtype=torch.float64) * 2 - 1) * lim
    return init


tf.contrib.layers.xavier_initializer = xavier_initializer
tf.contrib.layers.variance_scaling_initializer = variance_scaling_initializer
tf.initializers.lecun_uniform = lecun_uniform
tf.ones_initializer = lambda *a, **k: (lambda shape: torch.ones(shape, dtype=torch.float64))
tf.zeros_initializer = lambda *a, **k: (lambda shape: torch.zeros(shape, dtype=torch.float64))


def l2_regularizer(scale, scope=None):
    """tf.contrib.layers.l2_regularizer: scale * tf.nn.l2_loss(w) = scale * sum(w^2) / 2 (None when scale == 0)"""
    scale_v = float(scale)

    def reg(w):
        return scale_v * (w * w).sum() / 2.0
    return reg if scale_v != 0.0 else None


tf.contrib.layers.l2_regularizer = l2_regularizer


def get_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, **k):
    full = (_scope_name() + '/' + name) if S.scope else name
    if full in S.vars:
        return S.vars[full]               # AUTO_REUSE / reuse=True: the same variable (regulariser registered once)
    shape = _shape(shape)
    if full in S.preset:
        val = torch.as_tensor(np.asarray(S.preset[full]), dtype=S.float).reshape(shape).clone()
    else:
        init = initializer if initializer is not None else (S.scope_init[-1] or xavier_initializer())
        val = init(shape).to(S.float)
    v = val.detach().clone().requires_grad_(bool(trainable))
    v.op = types.SimpleNamespace(name=full)
    v.var_name = full
    S.vars[full] = v
    if regularizer is not None:
        S.regs[full] = regularizer
    return v


tf.get_variable = get_variable
tf.trainable_variables = lambda: list(S.vars.values())
tf.losses.get_regularization_loss = lambda *a, **k: sum((r(S.vars[n]) for n, r in S.regs.items()), torch.zeros((), dtype=S.float))


def placeholder(dtype=None, shape=None, name=None):
    v = S.feeds[name]
    if isinstance(v, list):
        v = v.pop(0)
    t = _t(v)
    if dtype in (S.float, torch.float32, torch.float64):
        t = t.to(S.float)
    elif dtype is not None and dtype != 'string':
        t = t.to(dtype)
    if shape is not None:
        assert list(t.shape) == [int(d) for d in shape], (name, t.shape, shape)
    return t


tf.placeholder = placeholder
tf.constant = lambda value, dtype=None, shape=None, name=None, **k: _t(value, dtype)
tf.Variable = lambda value, **k: _t(value)

# ---------------------------------------------------------------------------------------------- array ops
tf.shape = lambda x, name=None, **k: torch.tensor(list(_t(x).shape), dtype=torch.int32)
tf.python.ops.array_ops.shape = tf.shape
tf.size = lambda x, **k: torch.tensor(_t(x).numel(), dtype=torch.int32)
tf.rank = lambda x, **k: torch.tensor(_t(x).dim(), dtype=torch.int32)


def _to_float(x):
    """int -> tf.float32 rounds to float32 whatever precision the rest of the graph runs in (the reference casts int64
    millisecond timestamps to float32 BEFORE subtracting them, nar_model.py:1055-1060: that rounding is part of its result)"""
    x = _t(x)
    if not x.dtype.is_floating_point:
        x = x.to(torch.float32)
    return x.to(S.float)


tf.cast = lambda x, dtype, name=None: _to_float(x) if dtype in (torch.float32, S.float) else _t(x).to(dtype)
tf.to_float = lambda x, name=None: _to_float(x)
tf.to_int32 = lambda x, name=None: _t(x).to(torch.int32)
tf.to_int64 = lambda x, name=None: _t(x).to(torch.int64)
tf.sign = lambda x, name=None: torch.sign(_t(x))
tf.abs = lambda x, name=None: torch.abs(_t(x))
tf.expand_dims = lambda x, axis=None, name=None, dim=None: _t(x).unsqueeze((axis if axis is not None else dim)[0] if isinstance(axis if axis is not None else dim, (list, tuple)) else (axis if axis is not None else dim))
tf.squeeze = lambda x, axis=None, name=None: _t(x).squeeze() if axis is None else _t(x).squeeze(axis if not isinstance(axis, (list, tuple)) else axis[0])
tf.reshape = lambda x, shape, name=None: _t(x).reshape(_shape(shape))
tf.tile = lambda x, multiples, name=None: _t(x).repeat(*_shape(multiples))
tf.zeros = lambda shape, dtype=None, name=None: torch.zeros(_shape(shape), dtype=dtype or S.float)
tf.ones = lambda shape, dtype=None, name=None: torch.ones(_shape(shape), dtype=dtype or S.float)
tf.zeros_like = lambda x, dtype=None, name=None, **k: torch.zeros_like(_t(x), dtype=dtype)
tf.eye = lambda n, **k: torch.eye(_int(n), dtype=S.float)
tf.range = lambda *a, **k: torch.arange(*[_int(v) for v in a], dtype=k.get('dtype') or torch.int32)
tf.stack = lambda values, axis=0, name=None: torch.stack([_t(v) for v in values], dim=axis)
tf.slice = lambda x, begin, size, name=None: _t(x)[tuple(slice(b, None if s < 0 else b + s) for b, s in zip(_shape(begin), _shape(size)))]


def concat(values, axis, name=None):
    ts = [_t(v) for v in values]
    ts = [t for t in ts]
    return torch.cat(ts, dim=_int(axis))


tf.concat = concat
tf.boolean_mask = lambda tensor, mask, name=None, **k: _t(tensor)[_t(mask).to(torch.bool)]
tf.gather = lambda params, indices, name=None, **k: _t(params)[_t(indices).long()]
tf.nn.embedding_lookup = lambda params, ids, name=None, **k: _t(params)[_t(ids).long()]


def gather_nd(params, indices, name=None):
    idx = _t(indices).long()
    return _t(params)[tuple(idx[..., i] for i in range(idx.shape[-1]))]


tf.gather_nd = gather_nd


def one_hot(indices, depth, name=None, **k):
    idx = _t(indices).long()
    depth = _int(depth)
    out = torch.zeros(tuple(idx.shape) + (depth,), dtype=S.float)
    ok = (idx >= 0) & (idx < depth)                       # out-of-range indices give an all-zero row (TF semantics)
    out.scatter_(-1, idx.clamp(0, depth - 1).unsqueeze(-1), ok.to(S.float).unsqueeze(-1))
    return out


tf.one_hot = one_hot


def unique(x, out_idx=torch.int32, name=None):
    """tf.unique: values in order of FIRST occurrence, idx maps every element to its slot"""
    x = _t(x)
    vals, inv = torch.unique(x, return_inverse=True)          # sorted
    first = torch.full((vals.numel(),), x.numel(), dtype=torch.int64)
    first.scatter_reduce_(0, inv, torch.arange(x.numel()), reduce='amin')
    order = torch.argsort(first)                              # sorted slot -> rank by first occurrence
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel())
    Unique = collections.namedtuple('Unique', ['y', 'idx'])
    return Unique(vals[order], rank[inv].to(out_idx))


tf.unique = unique


def setdiff1d(x, y, index_dtype=torch.int32, name=None):
    """values of x (order and repetitions kept) that do not occur in y"""
    x, y = _t(x), _t(y)
    keep = ~torch.isin(x, y)
    idx = torch.nonzero(keep).reshape(-1)
    Out = collections.namedtuple('ListDiff', ['out', 'idx'])
    return Out(x[keep], idx.to(index_dtype))


tf.setdiff1d = setdiff1d


def unsorted_segment_min(data, segment_ids, num_segments, name=None):
    data, seg = _t(data), _t(segment_ids).long()
    n = _int(num_segments)
    out = torch.full((n,), torch.iinfo(data.dtype).max if not data.dtype.is_floating_point else float('inf'), dtype=data.dtype)
    out.scatter_reduce_(0, seg, data, reduce='amin')
    return out


tf.unsorted_segment_min = unsorted_segment_min


def where(condition, x=None, y=None, name=None):
    c = _t(condition).to(torch.bool)
    if x is None:
        return torch.nonzero(c)                               # int64 [N, rank], row-major
    return torch.where(c, _t(x), _t(y))


tf.where = where


def sequence_mask(lengths, maxlen=None, dtype=torch.bool, name=None):
    lengths = _t(lengths).long()
    m = _int(maxlen) if maxlen is not None else int(lengths.max().item())
    return (torch.arange(m)[None, :] < lengths[..., None]).to(dtype)


tf.sequence_mask = sequence_mask


def sparse_to_dense(sparse_indices, output_shape, sparse_values, default_value=0, **k):
    out = torch.full(_shape(output_shape), default_value, dtype=torch.int32)
    for i, v in zip(_shape(sparse_indices), _shape(sparse_values)):
        out[i] = v
    return out


tf.sparse_to_dense = sparse_to_dense


def dense_to_sparse(tensor, eos_token=0, **k):
    t = _t(tensor)
    idx = torch.nonzero(t != eos_token)
    return types.SimpleNamespace(indices=idx, values=t[tuple(idx.t())], dense_shape=torch.tensor(list(t.shape)))


tf.contrib.layers.dense_to_sparse = dense_to_sparse

# ---------------------------------------------------------------------------------------------- math


def _red(fn):
    def f(x, axis=None, keepdims=False, name=None, keep_dims=None, **k):
        x = _t(x)
        kd = bool(keepdims if keep_dims is None else keep_dims)
        if axis is None:
            r = fn(x)
            return r.reshape([1] * x.dim()) if kd else r
        ax = tuple(axis) if isinstance(axis, (list, tuple)) else (_int(axis),)
        for a in sorted((a % x.dim() for a in ax), reverse=True):
            x = fn(x, dim=a, keepdim=kd)
            x = x[0] if isinstance(x, tuple) else x
        return x
    return f


def _max(x, dim=None, keepdim=False):
    return x.max() if dim is None else x.max(dim=dim, keepdim=keepdim)[0]


def _min(x, dim=None, keepdim=False):
    return x.min() if dim is None else x.min(dim=dim, keepdim=keepdim)[0]


tf.reduce_sum = _red(torch.sum)
tf.reduce_mean = _red(torch.mean)
tf.reduce_max = _red(_max)
tf.reduce_min = _red(_min)
tf.log = lambda x, name=None: torch.log(_t(x))
tf.sqrt = lambda x, name=None: torch.sqrt(_t(x))
tf.pow = lambda x, y, name=None: torch.pow(_t(x), _t(y))
tf.maximum = lambda x, y, name=None: torch.maximum(_t(x), _t(y).to(_t(x).dtype))
tf.minimum = lambda x, y, name=None: torch.minimum(_t(x), _t(y).to(_t(x).dtype))
tf.multiply = lambda x, y, name=None: _t(x) * _t(y)
tf.subtract = lambda x, y, name=None: _t(x) - _t(y)
tf.add = lambda x, y, name=None: _t(x) + _t(y)
tf.div = lambda x, y, name=None: (_t(x) / _t(y)) if _t(x).dtype.is_floating_point else torch.div(_t(x), _t(y), rounding_mode='floor')
tf.mod = lambda x, y, name=None: torch.remainder(_t(x), _t(y))
tf.equal = lambda x, y, name=None: torch.eq(_t(x), _t(y))
tf.logical_and = lambda x, y, name=None: torch.logical_and(_t(x), _t(y))
tf.is_nan = lambda x, name=None: torch.isnan(_t(x))
tf.matmul = lambda a, b, transpose_a=False, transpose_b=False, name=None: (_t(a).transpose(-1, -2) if transpose_a else _t(a)) @ (_t(b).transpose(-1, -2) if transpose_b else _t(b))
tf.sigmoid = lambda x, name=None: torch.sigmoid(_t(x))
tf.tanh = lambda x, name=None: torch.tanh(_t(x))


def moments(x, axes, name=None, keep_dims=False, **k):
    x = _t(x)
    ax = tuple(_shape(axes))
    mean = x.mean(dim=ax, keepdim=keep_dims)
    var = ((x - x.mean(dim=ax, keepdim=True)) ** 2).mean(dim=ax, keepdim=keep_dims)      # biased, like tf.nn.moments
    return mean, var


tf.nn.moments = moments
tf.nn.relu = lambda x, name=None: torch.relu(_t(x))
tf.nn.leaky_relu = lambda x, alpha=0.2, name=None: torch.where(_t(x) > 0, _t(x), _t(x) * alpha)      # max(x, alpha*x)
tf.nn.tanh = lambda x, name=None: torch.tanh(_t(x))
tf.nn.sigmoid = lambda x, name=None: torch.sigmoid(_t(x))
tf.nn.l2_loss = lambda x, name=None: (_t(x) ** 2).sum() / 2.0
tf.nn.l2_normalize = lambda x, axis=None, epsilon=1e-12, name=None, dim=None: _t(x) / torch.sqrt(torch.clamp((_t(x) ** 2).sum(dim=axis if axis is not None else dim, keepdim=True), min=epsilon))
tf.nn.zero_fraction = lambda x, name=None: (_t(x) == 0).to(S.float).mean()


def softmax(logits, axis=-1, name=None, dim=None):
    x = _t(logits)
    S.softmax_inputs.append(x.detach().clone())
    return torch.softmax(x, dim=axis if dim is None else dim)


tf.nn.softmax = softmax


def top_k(x, k=1, sorted=True, name=None):
    x = _t(x)
    v, i = torch.sort(x, dim=-1, descending=True, stable=True)        # ties: lower index first (tf.nn.top_k)
    k = _int(k)
    TopK = collections.namedtuple('TopKV2', ['values', 'indices'])
    return TopK(v[..., :k], i[..., :k].to(torch.int32))


tf.nn.top_k = top_k

# ---------------------------------------------------------------------------------------------- control flow / random


def map_fn(fn, elems, dtype=None, **k):
    elems = _t(elems)
    outs = [fn(elems[i]) for i in range(elems.shape[0])]
    return torch.stack([_t(o) for o in outs]) if outs else torch.zeros((0,), dtype=elems.dtype)


tf.map_fn = map_fn
tf.cond = lambda pred, true_fn=None, false_fn=None, **k: (true_fn() if bool(_t(pred).item()) else false_fn())


def while_loop(cond, body, loop_vars, **k):
    lv = list(loop_vars)
    while bool(_t(cond(*lv)).item()):
        lv = list(body(*lv))
    return lv


tf.while_loop = while_loop


def random_shuffle(value, seed=None, name=None):
    v = _t(value)
    return v[torch.randperm(v.shape[0], generator=S.rng)]


tf.random_shuffle = random_shuffle
tf.random.shuffle = random_shuffle

# ---------------------------------------------------------------------------------------------- layers


class Dense:
    """tf.layers.Dense: variables are created at the FIRST call, under the variable scope active at that call"""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name=None, **k):
        self.units, self.activation, self.use_bias = int(units), activation, use_bias
        self.kernel_initializer, self.kernel_regularizer, self.bias_regularizer = kernel_initializer, kernel_regularizer, bias_regularizer
        self.name = name or 'dense'
        self.kernel = self.bias = None

    def __call__(self, inputs):
        x = _t(inputs)
        if self.kernel is None:
            with variable_scope(self.name):
                self.kernel = get_variable('kernel', [x.shape[-1], self.units], initializer=self.kernel_initializer,
                                           regularizer=self.kernel_regularizer)
                if self.use_bias:
                    self.bias = get_variable('bias', [self.units], initializer=tf.zeros_initializer(),
                                             regularizer=self.bias_regularizer)
        y = x.to(S.float) @ self.kernel
        if self.use_bias:
            y = y + self.bias
        return self.activation(y) if self.activation is not None else y


tf.layers.Dense = Dense
tf.layers.dense = lambda inputs, units, name=None, **k: Dense(units, name=name, **k)(inputs)


def dropout(inputs, rate=0.5, noise_shape=None, seed=None, training=False, name=None):
    x = _t(inputs)
    if not training or float(rate) == 0.0:
        return x
    keep = 1.0 - float(rate)
    mask = (torch.rand(x.shape, generator=S.rng, dtype=torch.float64) < keep)
    S.dropout_masks.append(mask.clone())
    return x * mask.to(x.dtype) / keep


tf.layers.dropout = dropout

# ---------------------------------------------------------------------------------------------- RNN (tf.contrib.rnn)


class UGRNNCell:
    """tf.contrib.rnn.UGRNNCell (TF 1.12 contrib/rnn/python/ops/rnn_cell.py): one Linear over concat(inputs, state) to
    2*units, split into (gate, candidate); g = sigmoid(gate + forget_bias), c = tanh(candidate);
    new_state = g * state + (1 - g) * c; output = new_state."""

    def __init__(self, num_units, initializer=None, forget_bias=1.0, activation=None, reuse=None):
        self.num_units, self.forget_bias = int(num_units), float(forget_bias)
        self.activation = activation or torch.tanh
        self.scope_name = 'ugrnn_cell'
        self.kernel = None

    @property
    def state_size(self):
        return self.num_units

    def __call__(self, inputs, state):
        if self.kernel is None:
            with variable_scope(self.scope_name):
                self.kernel = get_variable('kernel', [inputs.shape[-1] + self.num_units, 2 * self.num_units])
                self.bias = get_variable('bias', [2 * self.num_units], initializer=tf.zeros_initializer())
        z = torch.cat([inputs, state], dim=1) @ self.kernel + self.bias
        g_act, c_act = z[:, :self.num_units], z[:, self.num_units:]
        c = self.activation(c_act)
        g = torch.sigmoid(g_act + self.forget_bias)
        new = g * state + (1.0 - g) * c
        return new, new


class GRUCell:
    """tf.nn.rnn_cell.GRUCell: gates = sigmoid([x, h] Wg + bg) (bias init 1) -> r, u; c = tanh([x, r*h] Wc + bc);
    h' = u * h + (1 - u) * c."""

    def __init__(self, num_units, **k):
        self.num_units = int(num_units)
        self.scope_name = 'gru_cell'
        self.gk = None

    @property
    def state_size(self):
        return self.num_units

    def __call__(self, inputs, state):
        n = self.num_units
        if self.gk is None:
            with variable_scope(self.scope_name):
                with variable_scope('gates'):
                    self.gk = get_variable('kernel', [inputs.shape[-1] + n, 2 * n])
                    self.gb = get_variable('bias', [2 * n], initializer=tf.ones_initializer())
                with variable_scope('candidate'):
                    self.ck = get_variable('kernel', [inputs.shape[-1] + n, n])
                    self.cb = get_variable('bias', [n], initializer=tf.zeros_initializer())
        gates = torch.sigmoid(torch.cat([inputs, state], 1) @ self.gk + self.gb)
        r, u = gates[:, :n], gates[:, n:]
        c = torch.tanh(torch.cat([inputs, r * state], 1) @ self.ck + self.cb)
        new = u * state + (1 - u) * c
        return new, new


class DropoutWrapper:
    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, state_keep_prob=1.0, **k):
        self.cell, self.output_keep_prob = cell, float(output_keep_prob)

    @property
    def state_size(self):
        return self.cell.state_size

    def __call__(self, inputs, state):
        out, new = self.cell(inputs, state)
        if self.output_keep_prob < 1.0:
            out = dropout(out, rate=1.0 - self.output_keep_prob, training=True)
        return out, new


class MultiRNNCell:
    def __init__(self, cells, state_is_tuple=True):
        self.cells = list(cells)

    def __call__(self, inputs, states):
        new_states = []
        cur = inputs
        for i, cell in enumerate(self.cells):
            with variable_scope('cell_%d' % i):
                cur, ns = cell(cur, states[i])
            new_states.append(ns)
        return cur, tuple(new_states)


def dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, scope=None, **k):
    """tf.nn.dynamic_rnn (batch-major): past a sequence's length the output is zero and the state is carried over"""
    x = _t(inputs).to(S.float)
    B, T, _ = x.shape
    lengths = _t(sequence_length).long() if sequence_length is not None else torch.full((B,), T)
    cells = cell.cells if isinstance(cell, MultiRNNCell) else [cell]
    state = tuple(torch.zeros(B, c.state_size, dtype=S.float) for c in cells)
    outs = []
    with variable_scope(scope or 'rnn'):
        for t in range(T):
            if isinstance(cell, MultiRNNCell):
                with variable_scope('multi_rnn_cell'):
                    out, new_state = cell(x[:, t], state)
            else:
                out, ns = cell(x[:, t], state[0])
                new_state = (ns,)
            live = (t < lengths).to(S.float).unsqueeze(1)
            outs.append(out * live)
            state = tuple(live * n + (1 - live) * s for n, s in zip(new_state, state))
    final = state if isinstance(cell, MultiRNNCell) else state[0]
    return torch.stack(outs, dim=1), final


tf.contrib.rnn.UGRNNCell = UGRNNCell
tf.contrib.rnn.MultiRNNCell = MultiRNNCell
tf.nn.rnn_cell.GRUCell = GRUCell
tf.nn.rnn_cell.DropoutWrapper = DropoutWrapper
tf.nn.dynamic_rnn = dynamic_rnn

# ---------------------------------------------------------------------------------------------- optimiser / hooks


class AdamOptimizer:
    """tf.train.AdamOptimizer: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t); m, v moments; var -= lr_t * m / (sqrt(v) + eps)"""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), float(beta1), float(beta2), float(epsilon)

    def compute_gradients(self, loss, var_list=None, **k):
        names = [n for n, v in S.vars.items() if v.requires_grad]
        gs = torch.autograd.grad(loss, [S.vars[n] for n in names], allow_unused=True)
        S.grads = collections.OrderedDict((n, g) for n, g in zip(names, gs))
        return [(g, S.vars[n]) for n, g in zip(names, gs)]

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        t = 1
        lr_t = self.lr * np.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        S.vars_after = collections.OrderedDict()
        for g, v in grads_and_vars:
            if g is None:
                S.vars_after[v.var_name] = v.detach().clone()
                continue
            m = (1 - self.b1) * g
            vv = (1 - self.b2) * g * g
            S.vars_after[v.var_name] = (v - lr_t * m / (torch.sqrt(vv) + self.eps)).detach()
        return None


tf.train.AdamOptimizer = AdamOptimizer
tf.train.get_global_step = lambda *a, **k: None
tf.train.SessionRunHook = object
tf.train.SessionRunArgs = lambda *a, **k: types.SimpleNamespace(fetches=a, kwargs=k)

# ---------------------------------------------------------------------------------------------- streaming metrics (one batch)


def sparse_recall_at_top_k(labels, top_k_predictions, weights=None, name=None, **k):
    """batch value of tf.contrib.metrics.sparse_recall_at_top_k with one label per row: weighted hit rate"""
    lab, pred = _t(labels), _t(top_k_predictions)
    hit = (pred == lab).any(dim=-1).to(S.float)
    w = _t(weights).to(S.float) if weights is not None else torch.ones_like(hit)
    val = (hit * w).sum() / torch.clamp(w.sum(), min=1e-30)
    return val, val


def metrics_mean(values, weights=None, name=None, **k):
    v = _t(values).to(S.float)
    if weights is None:
        val = v.mean() if v.numel() else torch.zeros((), dtype=S.float)
    else:
        w = _t(weights).to(S.float)
        val = (v * w).sum() / torch.clamp(w.sum(), min=1e-30)
    return val, val


tf.contrib.metrics.sparse_recall_at_top_k = sparse_recall_at_top_k
tf.metrics.mean = metrics_mean
tf.contrib.lookup.HashTable = lambda *a, **k: None
tf.contrib.lookup.KeyValueTensorInitializer = lambda *a, **k: None

# ---------------------------------------------------------------------------------------------- tf.data / example parsing
# (for /root/reference/nar_module/nar/datasets.py: TFRecordDataset -> map(parse_sequence_example) -> padded_batch -> map)
tf.data = _mod('tensorflow.data')
tf.FixedLenFeature = collections.namedtuple('FixedLenFeature', ['shape', 'dtype'])
tf.FixedLenSequenceFeature = lambda shape, dtype, **k: types.SimpleNamespace(shape=shape, dtype=dtype)
tf.convert_to_tensor = lambda x, **k: _t(x)


def _np_dtype(dtype):
    return {torch.int64: np.int64, torch.int32: np.int32, torch.float32: np.float32, torch.float64: np.float64}[dtype]


def parse_single_sequence_example(serialized, context_features=None, sequence_features=None, example_name=None, name=None):
    """S.example_decoder(bytes) -> (context {name: values}, feature_lists {name: values per step}) does the protobuf part
    (the generator plugs in Google's protobuf runtime); FixedLenFeature([]) is a scalar, FixedLenSequenceFeature([]) a vector"""
    ctx, seqs = S.example_decoder(serialized)
    context_parsed = {k: torch.as_tensor(np.asarray(ctx[k], dtype=_np_dtype(spec.dtype)).reshape(()))
                      for k, spec in (context_features or {}).items()}
    sequence_parsed = {k: torch.as_tensor(np.asarray(seqs[k], dtype=_np_dtype(spec.dtype)).reshape(-1))
                       for k, spec in (sequence_features or {}).items()}
    return context_parsed, sequence_parsed


tf.parse_single_sequence_example = parse_single_sequence_example


class Dataset:
    def __init__(self, gen_fn):
        self._gen = gen_fn

    def __iter__(self):
        return iter(self._gen())

    def map(self, fn, num_parallel_calls=None):
        src = self._gen
        return Dataset(lambda: (fn(x) for x in src()))

    def prefetch(self, n):
        return self

    def padded_batch(self, batch_size, padded_shapes, padding_values=None, drop_remainder=False):
        """tf.data padded_batch: a [None] component is zero-padded to the longest of the batch, then everything is stacked"""
        src = self._gen

        def gen():
            it = iter(src())
            while True:
                chunk = []
                for x in it:
                    chunk.append(x)
                    if len(chunk) == batch_size:
                        break
                if not chunk or (drop_remainder and len(chunk) < batch_size):
                    return
                out = {}
                for k in chunk[0]:
                    vals = [_t(c[k]) for c in chunk]
                    n = max(v.shape[0] for v in vals)
                    out[k] = torch.stack([torch.cat([v, torch.zeros(n - v.shape[0], dtype=v.dtype)]) for v in vals])
                yield out
        return Dataset(gen)

    def make_one_shot_iterator(self):
        it = iter(self._gen())
        return types.SimpleNamespace(get_next=lambda: next(it))


def _read_tfrecords(path, gzip_compressed):
    import gzip
    import struct
    with (gzip.open(path, 'rb') if gzip_compressed else open(path, 'rb')) as f:
        while True:
            head = f.read(12)                         # uint64 length + masked crc32c of the length (not verified here)
            if len(head) < 12:
                return
            n = struct.unpack('<Q', head[:8])[0]
            data = f.read(n)
            f.read(4)                                 # masked crc32c of the data
            yield data


def TFRecordDataset(filenames, compression_type=None, **k):
    import glob as _glob
    names = [filenames] if isinstance(filenames, str) else list(filenames)
    files = []
    for n in names:
        files.extend(sorted(_glob.glob(n)) or [n])
    return Dataset(lambda: (r for p in files for r in _read_tfrecords(p, compression_type == 'GZIP')))


tf.data.TFRecordDataset = TFRecordDataset
tf.data.Dataset = Dataset

# ---------------------------------------------------------------------------------------------- tf.flags / EstimatorSpec
# (for /root/reference/nar_module/nar/nar_trainer_gcom.py: flag definitions at import time, nar_module_model_fn)
tf.flags = _mod('tensorflow.flags')
tf.flags.FLAGS = types.SimpleNamespace()
tf.app = _mod('tensorflow.app')
tf.app.flags = tf.flags
tf.app.run = lambda *a, **k: None


def _define(name, default=None, help=None, **k):
    setattr(tf.flags.FLAGS, name, default)


for _n in ('DEFINE_integer', 'DEFINE_float', 'DEFINE_string', 'DEFINE_boolean', 'DEFINE_bool', 'DEFINE_list'):
    setattr(tf.flags, _n, _define)


class EstimatorSpec:
    def __init__(self, mode, predictions=None, loss=None, train_op=None, eval_metric_ops=None, training_chief_hooks=None,
                 training_hooks=None, evaluation_hooks=None, **k):
        self.mode, self.loss, self.train_op, self.eval_metric_ops = mode, loss, train_op, eval_metric_ops
        self.training_chief_hooks, self.evaluation_hooks = training_chief_hooks, evaluation_hooks


tf.estimator.EstimatorSpec = EstimatorSpec
tf.estimator.Estimator = lambda *a, **k: types.SimpleNamespace(args=a, kwargs=k)
tf.estimator.RunConfig = lambda *a, **k: types.SimpleNamespace(kwargs=k)
tf.set_random_seed = lambda *a, **k: None
