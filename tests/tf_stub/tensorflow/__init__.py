"""TEST INFRASTRUCTURE ONLY -- a NumPy-backed, eager stand-in for the slice of the TensorFlow 1.13 API that the
reference's primitives use (models/common/{nade,rbm,dbn,model}.py, utils/{sequences,auxiliary}.py,
metrics/statistical.py: ~50 `tf.*` names), so that the REFERENCE'S OWN MODULES can be imported unmodified from
/root/reference and executed to generate golden vectors (tools/make_golden_ref.py -> tests/golden/ref_*.npz).

What this pins and what it does not: loop order, transposes / reshapes / unstack indices, where `safe_log` and the eps
sit, which biases are added where, the Gibbs-chain structure, the CD-k update formula and the flatten order all come
from the reference's code as written; the semantics of each individual op (matmul, sigmoid, log, tile, where, ...) are
NumPy's, in float64, following the TF 1.13 / TFP 0.6 documentation. TensorFlow itself cannot be
installed here (no cp312 wheel of 1.13.1, no network).

Graph-mode notions are collapsed to eager values: a "tensor" is an ndarray subclass, `tf.Variable` holds a value and
supports assign / assign_add, `tf.metrics.*` return (value of this batch, the same value) pairs, summaries are None,
`tf.while_loop` / `tf.cond` run as Python control flow. Randomness: initialisers draw from a module-level Generator
(`set_random_seed`); Bernoulli sampling lives in the tensorflow_probability stub and consumes an injected uniform stream.
"""
import builtins
import contextlib
import types

import numpy as np
from scipy.special import expit as _expit

float32 = np.float64        # every float tensor is computed in float64 (see the module docstring)
float64 = np.float64
int32 = np.int64
int64 = np.int64
bool = np.bool_             # noqa: A001 - mirrors tf.bool

_rng = np.random.default_rng(0)


def set_random_seed(seed):
    global _rng
    _rng = np.random.default_rng(seed)


class TensorShape(tuple):
    ndims = property(len)

    def as_list(self):
        return list(self)


class Tensor(np.ndarray):
    """ndarray whose `.shape` also answers `.ndims` / `.as_list()` like tf.TensorShape, with a TF-style `.name`."""
    name = 'tensor:0'

    def __array_finalize__(self, obj):
        self.name = getattr(obj, 'name', 'tensor:0')

    @property
    def shape(self):
        return TensorShape(np.ndarray.shape.__get__(self))

    def get_shape(self):
        return self.shape

    def assign(self, value):
        self[...] = np.asarray(value).reshape(np.ndarray.shape.__get__(self))
        return self

    def assign_add(self, value):
        self[...] = np.asarray(self) + np.asarray(value).reshape(np.ndarray.shape.__get__(self))
        return self

    def eval(self, *a, **k):
        return np.asarray(self)


def _t(x, dtype=None, name=None):
    a = np.asarray(x, dtype=dtype)
    if a.dtype == np.float32:
        a = a.astype(np.float64)
    out = np.array(a, copy=True).view(Tensor) if not isinstance(x, Tensor) else x
    if name is not None:
        out.name = _scoped(name) + ':0'
    return out


# ----------------------------------------------------------------------------- scopes / graph
_scopes = []


def _scoped(name):
    prefix = ''.join(_scopes)
    return prefix + name


@contextlib.contextmanager
def _scope(name):
    if name is None:
        name = ''
    if name and not name.endswith('/'):
        name += '/'
    # TF: a scope name ending in '/' is taken as an absolute name scope
    _scopes.append(name)
    try:
        yield name
    finally:
        _scopes.pop()


def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None, **kw):
    if isinstance(name_or_scope, str) and name_or_scope.endswith('/'):
        return _absolute_scope(name_or_scope)
    return _scope(name_or_scope if name_or_scope is not None else default_name)


def name_scope(name=None, default_name=None, values=None):
    if isinstance(name, str) and name.endswith('/'):
        return _absolute_scope(name)
    return _scope(name if name is not None else default_name)


@contextlib.contextmanager
def _absolute_scope(name):
    global _scopes
    saved = _scopes
    _scopes = [name]
    try:
        yield name
    finally:
        _scopes = saved


class _Graph:
    def get_name_scope(self):
        return ''.join(_scopes).rstrip('/')


def get_default_graph():
    return _Graph()


def reset_default_graph():
    del _scopes[:]
    _collections.clear()
    _default_names.clear()


_default_names = {}


def unique_default_name(name):
    """variable_scope(None, default_name=name): name, name_1, name_2, ... within the enclosing scope."""
    key = _scoped(name)
    n = _default_names.get(key, 0)
    _default_names[key] = n + 1
    return name if n == 0 else f'{name}_{n}'


_collections = {}


def add_to_collection(name, value):
    _collections.setdefault(name, []).append(value)


def get_collection(name):
    return list(_collections.get(name, []))


# ----------------------------------------------------------------------------- variables / initialisers
def Variable(initial_value, dtype=None, name='Variable', trainable=True):
    v = np.array(np.asarray(initial_value), dtype=np.float64, copy=True).view(Tensor)
    v.name = _scoped(name) + ':0'
    return v


def truncated_normal_initializer(mean=0.0, stddev=1.0, seed=None, dtype=None):
    def init(shape, dtype=None, partition_info=None):
        out = _rng.normal(mean, stddev, size=tuple(shape))
        bad = np.abs(out - mean) > 2 * stddev                     # TF re-draws values beyond two standard deviations
        while bad.any():
            out[bad] = _rng.normal(mean, stddev, size=int(bad.sum()))
            bad = np.abs(out - mean) > 2 * stddev
        return _t(out)
    return init


def _xavier_initializer(uniform=True, seed=None, dtype=None):
    def init(shape, dtype=None, partition_info=None):
        fan_in, fan_out = shape[-2], shape[-1]
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        return _t(_rng.uniform(-limit, limit, size=tuple(shape)))
    return init


def zeros(shape, dtype=None, name=None):
    return _t(np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=np.float64 if dtype in (None, float32) else dtype))


def ones(shape, dtype=None, name=None):
    return _t(np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=np.float64 if dtype in (None, float32) else dtype))


def zeros_like(x, dtype=None, name=None):
    return _t(np.zeros_like(np.asarray(x)))


def constant(value, dtype=None, shape=None, name=None):
    a = np.asarray(value, dtype=None if dtype is None else dtype)
    if shape is not None:
        a = np.broadcast_to(a, shape)
    return _t(a)


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


# ----------------------------------------------------------------------------- shape ops
def shape(x, name=None):
    return np.asarray(np.asarray(x).shape, dtype=np.int64)


def reshape(tensor, shape, name=None):
    return _t(np.reshape(np.asarray(tensor), tuple(int(s) for s in np.asarray(shape).reshape(-1))))


def transpose(a, perm=None, name=None):
    return _t(np.transpose(np.asarray(a), perm))


def tile(x, multiples, name=None):
    return _t(np.tile(np.asarray(x), tuple(int(m) for m in multiples)))


def unstack(value, num=None, axis=0, name=None):
    a = np.asarray(value)
    return [_t(np.take(a, i, axis=axis)) for i in range(a.shape[axis])]


def stack(values, axis=0, name=None):
    return _t(np.stack([np.asarray(v) for v in values], axis=axis))


def concat(values, axis, name=None):
    return _t(np.concatenate([np.atleast_1d(np.asarray(v)) for v in values], axis=axis))


def squeeze(x, axis=None, name=None, squeeze_dims=None):
    axis = axis if axis is not None else squeeze_dims
    return _t(np.squeeze(np.asarray(x), axis=None if axis is None else tuple(np.atleast_1d(axis))))


def expand_dims(x, axis, name=None):
    return _t(np.expand_dims(np.asarray(x), axis))


def pad(tensor, paddings, mode='CONSTANT', name=None, constant_values=0):
    return _t(np.pad(np.asarray(tensor), [tuple(p) for p in paddings], mode='constant', constant_values=constant_values))


def split(value, num_or_size_splits, axis=0, name=None):
    a = np.asarray(value)
    if isinstance(num_or_size_splits, int):
        return [_t(p) for p in np.split(a, num_or_size_splits, axis=axis)]
    return [_t(p) for p in np.split(a, np.cumsum(num_or_size_splits)[:-1], axis=axis)]


def gather_nd(params, indices, name=None):
    idx = np.asarray(indices)
    return _t(np.asarray(params)[tuple(idx[:, i] for i in range(idx.shape[1]))])


def sequence_mask(lengths, maxlen=None, dtype=None, name=None):
    lengths = np.asarray(lengths)
    maxlen = int(lengths.max()) if maxlen is None else int(maxlen)
    return _t(np.arange(maxlen)[None, :] < lengths[:, None])


def where(condition, x=None, y=None, name=None):
    c = np.asarray(condition)
    if x is None and y is None:
        return _t(np.argwhere(c))                                   # row-major order, like tf.where
    return _t(np.where(c, np.asarray(x), np.asarray(y)))


# ----------------------------------------------------------------------------- math
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = np.asarray(a), np.asarray(b)
    a = a.T if transpose_a else a
    b = b.T if transpose_b else b
    return _t(a @ b)


def sigmoid(x, name=None):
    return _t(_expit(np.asarray(x, dtype=np.float64)))


def tanh(x, name=None):
    return _t(np.tanh(np.asarray(x)))


def log(x, name=None):
    with np.errstate(divide='ignore'):
        return _t(np.log(np.asarray(x, dtype=np.float64)))


def exp(x, name=None):
    return _t(np.exp(np.asarray(x, dtype=np.float64)))


def subtract(x, y, name=None):
    return _t(np.asarray(x) - np.asarray(y), name=name)


def add(x, y, name=None):
    return _t(np.asarray(x) + np.asarray(y), name=name)


def multiply(x, y, name=None):
    return _t(np.asarray(x) * np.asarray(y), name=name)


def divide(x, y, name=None):
    return _t(np.asarray(x) / np.asarray(y), name=name)


def to_float(x, name=None):
    return _t(np.asarray(x).astype(np.float64))


def cast(x, dtype, name=None):
    return _t(np.asarray(x).astype(dtype))


def greater_equal(x, y, name=None):
    return _t(np.asarray(x) >= np.asarray(y))


def greater(x, y, name=None):
    return _t(np.asarray(x) > np.asarray(y))


def less(x, y, name=None):
    return _t(np.asarray(x) < np.asarray(y))


def equal(x, y, name=None):
    return _t(np.asarray(x) == np.asarray(y))


def _reduce(fn, x, axis, keepdims):
    return _t(fn(np.asarray(x), axis=None if axis is None else tuple(np.atleast_1d(axis)), keepdims=builtins.bool(keepdims)))


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
    return _reduce(np.sum, x, axis, keepdims or keep_dims)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
    return _reduce(np.mean, x, axis, keepdims or keep_dims)


def reduce_min(x, axis=None, keepdims=False, name=None):
    return _reduce(np.min, x, axis, keepdims)


def reduce_max(x, axis=None, keepdims=False, name=None):
    return _reduce(np.max, x, axis, keepdims)


def stop_gradient(x, name=None):
    return _t(np.array(np.asarray(x), copy=True), name=name)


def identity(x, name=None):
    return _t(x, name=name)


# ----------------------------------------------------------------------------- control flow (eager)
def while_loop(cond, body, loop_vars, **kw):
    vars_ = list(loop_vars)
    while np.all(np.asarray(cond(*vars_))):
        vars_ = list(body(*vars_))
    return vars_


class _LazyScalar:
    """tf.cond on the `is_train` placeholder whose branches are Python numbers (common/rnn.py:120, the dropout keep
    probability): resolved when used, so one constructed model serves is_train = False and True."""

    def __init__(self, pred, a, b):
        self._pred, self._a, self._b = pred, a, b

    def __float__(self):
        return float(self._a if np.all(np.asarray(self._pred)) else self._b)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype or np.float64)


def cond(pred, true_fn=None, false_fn=None, name=None):
    if getattr(pred, 'name', '').split(':')[0].endswith('is_train'):
        a, b = true_fn(), false_fn()
        if isinstance(a, (int, float)) and isinstance(b, (int, float)):
            return _LazyScalar(pred, a, b)
    return true_fn() if np.all(np.asarray(pred)) else false_fn()


def group(*ops, **kw):
    return list(ops)


# ----------------------------------------------------------------------------- losses / metrics / summaries / train
def _log_loss(labels, predictions, weights=1.0, epsilon=1e-7, scope=None, loss_collection=None, reduction='weighted_sum_by_nonzero_weights'):
    t, p = np.asarray(labels, dtype=np.float64), np.asarray(predictions, dtype=np.float64)
    losses_ = -t * np.log(p + epsilon) - (1 - t) * np.log(1 - p + epsilon)        # tf.losses.log_loss, TF 1.13
    if reduction == 'none':
        return _t(losses_ * weights)
    return _t(np.mean(losses_ * weights))


losses = types.SimpleNamespace(log_loss=_log_loss, Reduction=types.SimpleNamespace(NONE='none', MEAN='mean'))


def _metric_mean(values, weights=None, name=None, **kw):
    v = _t(np.mean(np.asarray(values, dtype=np.float64)))
    return v, v


def _counts(labels, predictions):
    t, p = np.asarray(labels) > 0.5, np.asarray(predictions) > 0.5
    return float((t & p).sum()), float((~t & p).sum()), float((t & ~p).sum())


def _metric_accuracy(labels, predictions, weights=None, name=None, **kw):
    v = _t(np.mean(np.asarray(labels) == np.asarray(predictions)))
    return v, v


def _metric_precision(labels, predictions, weights=None, name=None, **kw):
    tp, fp, _ = _counts(labels, predictions)
    v = _t(tp / (tp + fp) if tp + fp > 0 else 0.0)
    return v, v


def _metric_recall(labels, predictions, weights=None, name=None, **kw):
    tp, _, fn = _counts(labels, predictions)
    v = _t(tp / (tp + fn) if tp + fn > 0 else 0.0)
    return v, v


metrics = types.SimpleNamespace(mean=_metric_mean, accuracy=_metric_accuracy, precision=_metric_precision,
                                recall=_metric_recall)


def _aggregate_metric_map(names_to_tuples):
    return ({k: v[0] for k, v in names_to_tuples.items()}, {k: v[1] for k, v in names_to_tuples.items()})


class _Anything:
    """Attribute sink for API corners the primitives only touch at import or for bookkeeping (tf.contrib.rnn, tf.nn, ...)."""

    def __init__(self, path):
        self._path = path

    def __getattr__(self, name):
        return _Anything(f'{self._path}.{name}')

    def __call__(self, *a, **k):
        raise NotImplementedError(f'{self._path} is not part of the NumPy TF stub (tests/tf_stub)')


contrib = types.SimpleNamespace(
    layers=types.SimpleNamespace(xavier_initializer=_xavier_initializer),
    metrics=types.SimpleNamespace(aggregate_metric_map=_aggregate_metric_map),
    rnn=_Anything('tf.contrib.rnn'), cudnn_rnn=_Anything('tf.contrib.cudnn_rnn'), seq2seq=_Anything('tf.contrib.seq2seq'))
nn = _Anything('tf.nn')
layers = _Anything('tf.layers')


class Summary:
    pass


class _MergedSummary(dict):
    """tf.summary.merge's result. Real TF returns a string tensor; multinn_core.py:242 then does
    `self._summaries['weights'] = ...`, which works when `_summaries` is the dict `_combine_track_metrics` builds
    (per-track modes) and raises TypeError when it is a merged tensor (Joint mode, multinn_joint.py:177-186 -- the
    reference's Joint build() cannot finish at HEAD). The stand-in accepts the assignment so that the Joint graph's
    numerics can still be recorded; summaries carry no numbers here."""


summary = types.SimpleNamespace(merge=lambda inputs, **k: _MergedSummary(), histogram=lambda *a, **k: None,
                                scalar=lambda *a, **k: None, FileWriter=_Anything('tf.summary.FileWriter'))


class _Saver:
    def __init__(self, *a, **k):
        pass

    def save(self, *a, **k):
        raise NotImplementedError('tf.train.Saver.save: not part of the NumPy TF stub')

    restore = save


train = types.SimpleNamespace(Saver=_Saver, get_checkpoint_state=lambda *a, **k: None,
                              AdamOptimizer=_Anything('tf.train.AdamOptimizer'),
                              GradientDescentOptimizer=_Anything('tf.train.GradientDescentOptimizer'))
Saver = _Saver


class Session:
    def __init__(self, *a, **k):
        pass

    def run(self, fetches, feed_dict=None):
        return fetches


__version__ = '1.13.1-numpy-stub'


# ================================================================================================ graph-level pieces
# (round 2, second step) enough of the graph-building API for the reference's MODE classes to run eagerly: placeholders
# that hold fed values, the LSTM cell stack and the seq2seq decoder loop, tf.scan. With these,
# tools/make_golden_ref.py builds `MultINN(config, params, 'composer' | 'jamming')` from the reference's own
# multinn/models/multinn/*.py and generators/*.py: padding, unstacking / stacking of tracks, the input / target shift,
# the flatten order, the Dense bias split and the loss come from the reference's code. The cell arithmetic below is
# TF 1.13's LSTMBlockCell as documented: gates i, ci, f, o = split([x, h] . kernel + bias);
# cs = tanh(ci) * sigmoid(i) + cs_prev * sigmoid(f + forget_bias); h = tanh(cs) * sigmoid(o); forget_bias = 0 for
# CudnnCompatibleLSTMCell.
class Dim(int):
    value = property(lambda s: int(s))


class _DimShape(TensorShape):
    def __getitem__(self, i):
        v = tuple.__getitem__(self, i)
        return _DimShape(v) if isinstance(i, slice) else Dim(v)


Tensor.shape = property(lambda self: _DimShape(np.ndarray.shape.__get__(self)))

_feeds = {}


def feed(**values):
    """Values of the placeholders created from now on, by name ('x', 'lengths', 'is_train')."""
    _feeds.update(values)


def placeholder(dtype, shape=None, name=None):
    if name not in _feeds:
        raise KeyError(f'tf stub: no value fed for placeholder {name!r} (tf.feed(name=value) before building the graph)')
    return _t(_feeds[name], name=name)


_variables = []
_Variable_plain = Variable


def Variable(initial_value, dtype=None, name='Variable', trainable=True):   # noqa: F811 - registers the variable
    v = _Variable_plain(initial_value, dtype, name, trainable)
    _variables.append(v)
    return v


def local_variables():
    return []


def global_variables():
    return list(_variables)


def variables_initializer(var_list, name='init'):
    return None


LSTMStateTuple = __import__('collections').namedtuple('LSTMStateTuple', ('c', 'h'))


class _LSTMBlockCell:
    def __init__(self, num_units, forget_bias=0.0, name='cudnn_compatible_lstm_cell'):
        self._num_units, self._forget_bias, self._name = num_units, forget_bias, name
        self.kernel = self.bias = None
        self._scope = None

    state_size = property(lambda s: LSTMStateTuple(s._num_units, s._num_units))
    output_size = property(lambda s: s._num_units)

    def zero_state(self, batch_size, dtype=None):
        z = lambda: _t(np.zeros((int(batch_size), self._num_units)))
        return LSTMStateTuple(z(), z())

    get_initial_state = lambda self, inputs=None, batch_size=None, dtype=None: self.zero_state(batch_size, dtype)

    @property
    def trainable_variables(self):
        return [v for v in (self.kernel, self.bias) if v is not None]

    variables = trainable_variables

    def __call__(self, inputs, state, scope=None):
        x, (cs_prev, h_prev) = np.asarray(inputs), state
        if self.kernel is None:
            with variable_scope(self._scope or self._name):
                init = contrib.layers.xavier_initializer()          # the variable scope's default: glorot_uniform
                self.kernel = Variable(init([x.shape[1] + self._num_units, 4 * self._num_units]), name='kernel')
                self.bias = Variable(np.zeros(4 * self._num_units), name='bias')
        xh = np.concatenate([x, np.asarray(h_prev)], axis=1)
        g = xh @ np.asarray(self.kernel) + np.asarray(self.bias)
        i, ci, f, o = np.split(g, 4, axis=1)
        cs = np.tanh(ci) * _expit(i) + np.asarray(cs_prev) * _expit(f + self._forget_bias)
        h = np.tanh(cs) * _expit(o)
        return _t(h), LSTMStateTuple(_t(cs), _t(h))


_dropout_uniforms = []


def push_dropout_uniforms(arrays):
    """Uniforms consumed by DropoutWrapper, one array per (step, layer) call in call order (tf.nn.dropout:
    x / keep * floor(keep + u))."""
    _dropout_uniforms.extend(np.asarray(a, dtype=np.float64) for a in arrays)


dropout_log = []
_drop_rng = [np.random.default_rng(0)]


def seed_dropout(seed):
    _drop_rng[0] = np.random.default_rng(seed)
    del dropout_log[:]


class _DropoutWrapper:
    def __init__(self, cell, input_keep_prob=1.0, output_keep_prob=1.0, **kw):
        self._cell, self._keep = cell, output_keep_prob

    state_size = property(lambda s: s._cell.state_size)
    output_size = property(lambda s: s._cell.output_size)
    trainable_variables = property(lambda s: s._cell.trainable_variables)
    variables = trainable_variables
    zero_state = lambda self, batch_size, dtype=None: self._cell.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        out, new_state = self._cell(inputs, state)
        keep = float(np.asarray(self._keep))
        if keep < 1.0:
            if _dropout_uniforms:
                u = _dropout_uniforms.pop(0)
            else:                                   # seeded draw, float32-representable, logged for the fixture
                u = _drop_rng[0].random(np.asarray(out).shape, dtype=np.float32).astype(np.float64)
            dropout_log.append(u)
            out = _t(np.asarray(out) / keep * np.floor(keep + u))
        return out, new_state


class _MultiRNNCell:
    def __init__(self, cells, state_is_tuple=True):
        self._cells = list(cells)
        for i, c in enumerate(self._cells):
            inner = getattr(c, '_cell', c)
            inner._scope = f'multi_rnn_cell/cell_{i}/{inner._name}'

    state_size = property(lambda s: tuple(c.state_size for c in s._cells))
    output_size = property(lambda s: s._cells[-1].output_size)

    @property
    def trainable_variables(self):
        return [v for c in self._cells for v in c.trainable_variables]

    variables = trainable_variables

    def zero_state(self, batch_size, dtype=None):
        return tuple(c.zero_state(batch_size, dtype) for c in self._cells)

    get_initial_state = lambda self, inputs=None, batch_size=None, dtype=None: self.zero_state(batch_size, dtype)

    def __call__(self, inputs, state, scope=None):
        new, x = [], inputs
        for c, s in zip(self._cells, state):
            x, ns = c(x, s)
            new.append(ns)
        return x, tuple(new)


class _TrainingHelper:
    def __init__(self, inputs, sequence_length, time_major=False, name=None):
        self.inputs, self.sequence_length = np.asarray(inputs), np.asarray(sequence_length)


class _BasicDecoder:
    def __init__(self, cell, helper, initial_state, output_layer=None):
        self.cell, self.helper, self.initial_state, self.output_layer = cell, helper, initial_state, output_layer


_DecoderOutput = __import__('collections').namedtuple('BasicDecoderOutput', ('rnn_output', 'sample_id'))


def _dynamic_decode(decoder, output_time_major=False, impute_finished=False, maximum_iterations=None, **kw):
    """tf.contrib.seq2seq.dynamic_decode with a TrainingHelper, impute_finished=False: the cell and the
    output layer run on every step t < max(sequence_length) for EVERY row; outputs past a row's length are not zeroed and
    its state keeps advancing. Returns (BasicDecoderOutput(rnn_output[B, T, C], sample_id), final_state, lengths)."""
    h = decoder.helper
    T = int(h.sequence_length.max())
    state, outs = decoder.initial_state, []
    for t in range(T):
        out, state = decoder.cell(_t(h.inputs[:, t]), state)
        if decoder.output_layer is not None:
            out = decoder.output_layer(out)
        outs.append(np.asarray(out))
    rnn_output = _t(np.stack(outs, axis=1))
    return _DecoderOutput(rnn_output, _t(np.argmax(np.asarray(rnn_output), -1))), state, _t(h.sequence_length)


def scan(fn, elems, initializer=None, **kw):
    """tf.scan over the leading axis; returns the stacked accumulator structure (tuples / namedtuples / lists of tensors)."""
    import collections.abc as _abc
    acc, outs = initializer, []
    n = np.asarray(elems).shape[0]
    for i in range(n):
        acc = fn(acc, _t(np.asarray(elems)[i]))
        outs.append(acc)

    def stack_struct(items):
        first = items[0]
        if isinstance(first, tuple) and hasattr(first, '_fields'):
            return type(first)(*(stack_struct([it[k] for it in items]) for k in range(len(first))))
        if isinstance(first, (tuple, list)):
            return type(first)(stack_struct([it[k] for it in items]) for k in range(len(first)))
        if first is None:
            return None
        return _t(np.stack([np.asarray(it) for it in items]))
    return stack_struct(outs)


def clip_by_global_norm(t_list, clip_norm, use_norm=None, name=None):
    gn = np.sqrt(sum(float((np.asarray(t) ** 2).sum()) for t in t_list))
    return [_t(np.asarray(t) * clip_norm / max(gn, clip_norm)) for t in t_list], _t(gn)


contrib.cudnn_rnn = types.SimpleNamespace(CudnnCompatibleLSTMCell=lambda num_units, reuse=None: _LSTMBlockCell(num_units, 0.0))
contrib.rnn = types.SimpleNamespace(MultiRNNCell=_MultiRNNCell, LSTMStateTuple=LSTMStateTuple,
                                    AttentionCellWrapper=_Anything('tf.contrib.rnn.AttentionCellWrapper'))
contrib.seq2seq = types.SimpleNamespace(TrainingHelper=_TrainingHelper, BasicDecoder=_BasicDecoder,
                                        dynamic_decode=_dynamic_decode)
nn = types.SimpleNamespace(rnn_cell=types.SimpleNamespace(DropoutWrapper=_DropoutWrapper, MultiRNNCell=_MultiRNNCell,
                                                          LSTMStateTuple=LSTMStateTuple),
                           dynamic_rnn=_Anything('tf.nn.dynamic_rnn'), sigmoid=sigmoid)


def _dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, **kw):
    """tf.nn.dynamic_rnn, batch-major: for rows with t >= sequence_length the output is zero and the state is copied
    through (frozen at the row's last valid step) -- unlike dynamic_decode(impute_finished=False) above."""
    x = np.asarray(inputs)
    B, T = x.shape[:2]
    lengths = np.full(B, T) if sequence_length is None else np.asarray(sequence_length).astype(np.int64)
    state = initial_state if initial_state is not None else cell.zero_state(B, dtype)
    outs = []

    def merge(new, old, live):
        if isinstance(new, tuple) and hasattr(new, '_fields'):
            return type(new)(*(merge(n, o, live) for n, o in zip(new, old)))
        if isinstance(new, (tuple, list)):
            return type(new)(merge(n, o, live) for n, o in zip(new, old))
        return _t(np.where(live[:, None], np.asarray(new), np.asarray(old)))
    for t in range(T):
        out, new_state = cell(_t(x[:, t]), state)
        live = t < lengths
        outs.append(np.where(live[:, None], np.asarray(out), 0.0))
        state = merge(new_state, state, live)
    return _t(np.stack(outs, axis=1)), state


nn.dynamic_rnn = _dynamic_rnn


class _Layers:
    @property
    def Dense(self):
        from tensorflow.python.layers.core import Dense
        return Dense


layers = _Layers()


# ================================================================================================ metrics/musical_tf.py
# (round 2, third step) the few extra ops `get_metric_summary_ops` uses; tf.summary.scalar records (scoped tag, value) so
# that tools/make_golden_musical.py can pin the TensorBoard tags and values of the sample scores.
def count_nonzero(x, axis=None, keepdims=False, dtype=None, name=None):
    return _t(np.count_nonzero(np.asarray(x), axis=axis, keepdims=keepdims).astype(np.float64 if dtype is not None else np.int64))


def reduce_any(x, axis=None, keepdims=False, name=None):
    return _t(np.any(np.asarray(x), axis=axis, keepdims=keepdims))


def reduce_all(x, axis=None, keepdims=False, name=None):
    return _t(np.all(np.asarray(x), axis=axis, keepdims=keepdims))


class _Eager:
    def __init__(self, a):
        self._a = np.asarray(a)

    def numpy(self):
        return self._a


def py_function(func, inp, Tout, name=None):
    return _t(np.asarray(func(*[_Eager(a) for a in inp]), dtype=np.float64))


scalar_log = []


def _summary_scalar(name, tensor, **kw):
    scalar_log.append((_scoped(name), float(np.asarray(tensor))))
    return None


summary.scalar = _summary_scalar
Summary = _Anything('tf.Summary')
