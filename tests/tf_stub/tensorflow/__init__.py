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


def cond(pred, true_fn=None, false_fn=None, name=None):
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


summary = types.SimpleNamespace(merge=lambda inputs, **k: None, histogram=lambda *a, **k: None,
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
