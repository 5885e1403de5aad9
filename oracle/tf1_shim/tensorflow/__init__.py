"""Minimal eager TF-1.x API shim on PyTorch fp64 -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's UNMODIFIED model-building Python
(`DMT_code/model/net/mmoe_transformer_unbias.py`, `TransformerModel*.py`, `base.py`,
`model/inference_mlp.py`) in this container, where `tensorflow==1.12` cannot be installed,
so that everything the reference decides in *Python* -- variable-scope nesting and reuse
(the shared encoder/decoder feed-forward), auto-numbered `dense`, `dense_1`, ... names,
concat order, the zero-pad row offset, mask construction, loss wiring -- is taken from the
reference itself rather than from a restatement.  Only the TF *primitives* are restated
here, each from its documented semantics.  `tests/golden/make_golden.py` puts this directory
on sys.path as `tensorflow`, imports the reference modules from /root/reference, runs them
on seeded inputs and commits the resulting vectors.

Everything is eager: a "tensor" is a torch.Tensor subclass that adds the handful of
tf.Tensor methods the reference uses (`get_shape().as_list()`, `shape[i].value`) and makes
augmented assignment (`x *= s`) out-of-place, as it is in a TF graph.  tf.float32 maps to
torch.float64 so the recorded vectors are exact-math references.

Variable scoping follows tensorflow/python/ops/variable_scope.py (1.12): string scopes
extend the current name, `reuse` is inherited, leaving a scope resets the default-name
counters of its sub-scopes (this is what makes the decoder's `ff()` land on the encoder's
`positionwise_feedforward/dense` again), `default_name` is uniquified with `_<n>`.
"""
import contextlib
import math

import torch

__version__ = "1.12.0-shim"

float32 = torch.float64
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool
string = "string"
AUTO_REUSE = "AUTO_REUSE"


# ----------------------------------------------------------------------------- tensors
class _Dim(int):
    @property
    def value(self):
        return int(self)


class _Shape(object):
    def __init__(self, dims):
        self._d = [_Dim(d) for d in dims]

    def as_list(self):
        return [int(d) for d in self._d]

    def __getitem__(self, i):
        r = self._d[i]
        return _Shape(r) if isinstance(i, slice) else r

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(self._d)


class Tensor(torch.Tensor):
    def get_shape(self):
        return _Shape(torch.Tensor.size(self))

    # tf.Tensor.__mul__ etc. run convert_to_tensor on Python lists (`mask * [1.0, 15.0, ...]`)
    @staticmethod
    def _c(o):
        return torch.as_tensor(o, dtype=torch.float64) if isinstance(o, (list, tuple)) else o

    def __mul__(self, o):
        return torch.Tensor.__mul__(self, Tensor._c(o))

    def __rmul__(self, o):
        return torch.Tensor.__rmul__(self, Tensor._c(o))

    def __add__(self, o):
        return torch.Tensor.__add__(self, Tensor._c(o))

    def __radd__(self, o):
        return torch.Tensor.__radd__(self, Tensor._c(o))

    def __sub__(self, o):
        return torch.Tensor.__sub__(self, Tensor._c(o))

    def __truediv__(self, o):
        return torch.Tensor.__truediv__(self, Tensor._c(o))

    # graph semantics: `x *= s` rebinds the Python name, it never mutates the producer's value
    def __imul__(self, o):
        return self * o

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __itruediv__(self, o):
        return self / o


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        y = x
    else:
        y = torch.as_tensor(x)
        if y.is_floating_point():
            y = y.to(torch.float64)
    if dtype is not None and y.dtype != dtype:
        y = y.to(dtype)
    return y.as_subclass(Tensor)


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices = _t(indices, torch.int64)
        self.values = _t(values)
        self.dense_shape = [int(d) for d in (dense_shape.tolist() if isinstance(dense_shape, torch.Tensor)
                                             else dense_shape)]


# ----------------------------------------------------------------------------- variable scopes
class _Store(object):
    def __init__(self):
        self.reset()

    def reset(self, seed=0):
        self.vars = {}
        self.order = []
        self.counts = {}
        self.scope = ("", None)    # (name, reuse)
        self.gen = torch.Generator().manual_seed(seed)
        self.trainable = []
        self.reg_losses = []


_S = _Store()


def reset_default_graph(seed=0):
    _S.reset(seed)


def global_variables():
    return [(n, _S.vars[n]) for n in _S.order]


def trainable_variables():
    return [(n, _S.vars[n]) for n in _S.order if n in _S.trainable]


class _VarScope(object):
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse

    def reuse_variables(self):
        self.reuse = True
        _S.scope = (self.name, True)


def get_variable_scope():
    return _VarScope(*_S.scope)


def _unique(prefix):
    cur = _S.scope[0]
    name = cur + "/" + prefix if cur else prefix
    if _S.counts.get(name, 0) == 0:
        return prefix
    i = 1
    while _S.counts.get("%s_%d" % (name, i), 0) > 0:
        i += 1
    return "%s_%d" % (prefix, i)


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, reuse=None, **_kw):
    old = _S.scope
    if isinstance(name_or_scope, _VarScope):
        # re-entering a captured scope object (tf.layers does this): jump to that name
        saved = dict(_S.counts)
        _S.counts[name_or_scope.name] = _S.counts.get(name_or_scope.name, 0) + 1
        _S.scope = (name_or_scope.name, reuse if reuse is not None else name_or_scope.reuse)
        try:
            yield _VarScope(*_S.scope)
        finally:
            _S.counts = saved
            _S.scope = old
        return
    piece = name_or_scope if name_or_scope is not None else _unique(default_name)
    new_name = old[0] + "/" + piece if old[0] else piece
    new_reuse = reuse or old[1]                      # re-using is inherited by sub-scopes
    _S.counts[new_name] = _S.counts.get(new_name, 0) + 1
    _S.scope = (new_name, new_reuse)
    try:
        yield _VarScope(new_name, new_reuse)
    finally:
        for k in list(_S.counts):                    # close_variable_subscopes
            if k.startswith(new_name + "/"):
                _S.counts[k] = 0
        _S.scope = old


@contextlib.contextmanager
def name_scope(name=None, *a, **k):
    yield name


@contextlib.contextmanager
def device(name=None):
    yield


@contextlib.contextmanager
def control_dependencies(deps=None):
    yield


def get_variable(name, shape=None, initializer=None, regularizer=None, trainable=True, dtype=None, **_kw):
    scope, reuse = _S.scope
    full = scope + "/" + name if scope else name
    if isinstance(shape, _Shape):
        shape = shape.as_list()
    elif isinstance(shape, int):
        shape = [shape]
    shape = [int(s) for s in shape] if shape is not None else None
    if full in _S.vars:
        if not reuse:
            raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True or "
                             "reuse=tf.AUTO_REUSE in VarScope?" % full)
        v = _S.vars[full]
        if shape is not None and list(v.shape) != shape:
            raise ValueError("Trying to share variable %s, but specified shape %s and found shape %s."
                             % (full, shape, list(v.shape)))
        return v
    if reuse is True:
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()." % full)
    init = initializer if initializer is not None else glorot_uniform_initializer()
    v = _t(init(shape)).clone().as_subclass(Tensor)
    v.requires_grad_(False)
    _S.vars[full] = v
    _S.order.append(full)
    if trainable:
        _S.trainable.append(full)
    if regularizer is not None:
        _S.reg_losses.append(regularizer(v))
    return v


def Variable(initial_value, name=None, trainable=True, **_kw):
    v = _t(initial_value)
    _S.vars[name or "Variable"] = v
    _S.order.append(name or "Variable")
    return v


# ----------------------------------------------------------------------------- initializers
def zeros_initializer():
    return lambda shape: torch.zeros(shape, dtype=torch.float64)


def ones_initializer():
    return lambda shape: torch.ones(shape, dtype=torch.float64)


def constant_initializer(value=0.0):
    return lambda shape: torch.full(shape, float(value), dtype=torch.float64)


def truncated_normal_initializer(mean=0.0, stddev=1.0, **_kw):
    def init(shape):
        t = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(t, mean, stddev, mean - 2 * stddev, mean + 2 * stddev, generator=_S.gen)
        return t
    return init


def glorot_uniform_initializer(**_kw):
    def init(shape):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, dtype=torch.float64, generator=_S.gen) * 2 - 1) * lim
    return init


# ----------------------------------------------------------------------------- ops
def constant(value, dtype=None, **_kw):
    return _t(value, dtype)


def convert_to_tensor(value, dtype=None, **_kw):
    # the one place the reference turns a numpy fp64 array into a tf.float32 constant (the sinusoid position table,
    # TransformerModel_util.py:262): the shim computes in fp64 but keeps the fp32 rounding of that constant
    import numpy as _np
    if dtype is not None and isinstance(value, _np.ndarray) and value.dtype == _np.float64:
        value = value.astype(_np.float32).astype(_np.float64)
    return _t(value, dtype)


def identity(x, name=None):
    return _t(x)


def ones(shape, dtype=float32):
    if isinstance(shape, torch.Tensor):
        shape = [int(shape)] if shape.dim() == 0 else shape.tolist()
    elif isinstance(shape, int):
        shape = [shape]
    return _t(torch.ones([int(s) for s in shape], dtype=dtype))


def zeros(shape, dtype=float32):
    if isinstance(shape, torch.Tensor):
        shape = [int(shape)] if shape.dim() == 0 else shape.tolist()
    return _t(torch.zeros([int(s) for s in shape], dtype=dtype))


def ones_like(x, dtype=None):
    return _t(torch.ones_like(_t(x), dtype=dtype))


def zeros_like(x, dtype=None):
    return _t(torch.zeros_like(_t(x), dtype=dtype))


def size(x):
    return int(_t(x).numel())


def shape(x):
    return [int(s) for s in _t(x).shape]


def cast(x, dtype):
    x = _t(x)
    if dtype in (torch.int32, torch.int64) and x.is_floating_point():
        # C float->int conversion truncates; non-finite input is undefined in TF, saturate here
        x = torch.nan_to_num(x, nan=0.0, posinf=2.0 ** 31 - 1, neginf=-2.0 ** 31).trunc()
    return _t(x.to(dtype))


def to_float(x):
    return cast(x, float32)


def to_int64(x):
    return cast(x, int64)


def range(*a, **k):   # noqa: A001 (tf.range)
    return _t(torch.arange(*[int(v) for v in a]))


def tile(x, multiples):
    return _t(_t(x).repeat(*[int(m) for m in multiples]))


def expand_dims(x, axis):
    return _t(_t(x).unsqueeze(axis))


def squeeze(x, axis=None):
    x = _t(x)
    return _t(x.squeeze() if axis is None else x.squeeze(axis))


def concat(values, axis):
    return _t(torch.cat([_t(v) for v in values], dim=axis))


def split(value, num_or_size_splits, axis=0):
    return [_t(p) for p in torch.chunk(_t(value), num_or_size_splits, dim=axis)]


def stack(values, axis=0):
    return _t(torch.stack([_t(v) for v in values], dim=axis))


def transpose(x, perm=None):
    x = _t(x)
    return _t(x.t() if perm is None else x.permute(*perm))


def reshape(x, shape):
    return _t(_t(x).reshape([int(s) for s in shape]))


def matmul(a, b):
    return _t(torch.matmul(_t(a), _t(b)))


def add(a, b):
    return _t(_t(a) + _t(b))


def divide(a, b):
    return _t(_t(a) / _t(b))


div = divide


def exp(x):
    return _t(torch.exp(_t(x)))


def log(x):
    return _t(torch.log(_t(x)))


def abs(x):   # noqa: A001
    return _t(torch.abs(_t(x)))


def sign(x):
    return _t(torch.sign(_t(x)))


def sigmoid(x):
    return _t(torch.sigmoid(_t(x)))


def equal(a, b):
    return _t(torch.eq(_t(a), _t(b)))


def greater(a, b):
    return _t(torch.gt(_t(a), _t(b)))


def where(cond, x=None, y=None):
    return _t(torch.where(_t(cond).to(torch.bool), _t(x), _t(y)))


def reduce_sum(x, axis=None, keep_dims=False, keepdims=False):
    x = _t(x)
    kd = keep_dims or keepdims
    return _t(x.sum() if axis is None else x.sum(dim=axis, keepdim=kd))


def reduce_mean(x, axis=None, keep_dims=False, keepdims=False):
    x = _t(x)
    kd = keep_dims or keepdims
    return _t(x.mean() if axis is None else x.mean(dim=axis, keepdim=kd))


def norm(x, ord=2, axis=None):   # noqa: A002
    return _t(torch.linalg.vector_norm(_t(x), ord=ord, dim=axis))


def clip_by_value(x, clip_value_min, clip_value_max):
    x = _t(x)
    lo, hi = _t(clip_value_min).to(x.dtype), _t(clip_value_max).to(x.dtype)
    return _t(torch.minimum(torch.maximum(x, lo), hi))


def sequence_mask(lengths, maxlen=None, dtype=bool):
    lengths = _t(lengths).long()
    maxlen = int(maxlen) if maxlen is not None else int(lengths.max())
    return _t((torch.arange(maxlen)[None, :] < lengths[:, None]).to(dtype))


def unique(x):
    vals, inv = torch.unique(_t(x), return_inverse=True)
    return _t(vals), _t(inv)


def gather(params, indices):
    return _t(_t(params)[_t(indices).long()])


def Print(x, *a, **k):
    return x


class _NN(object):
    @staticmethod
    def relu(x):
        return _t(torch.relu(_t(x)))

    @staticmethod
    def softmax(x, axis=-1):
        return _t(torch.softmax(_t(x), dim=axis))

    @staticmethod
    def embedding_lookup(params, ids):
        return _t(_t(params)[_t(ids).long()])

    @staticmethod
    def embedding_lookup_sparse(params, sp_ids, sp_weights, combiner="mean"):
        """segment-wise sum_j w_j * params[id_j] / sum_j w_j (combiner='mean')."""
        params = _t(params)
        seg = sp_ids.indices[:, 0].long()
        rows = params[sp_ids.values.long()]
        n = int(seg.max()) + 1 if seg.numel() else 0
        w = torch.ones(rows.shape[0], dtype=params.dtype) if sp_weights is None else sp_weights.values.to(params.dtype)
        num = torch.zeros(n, params.shape[1], dtype=params.dtype).index_add_(0, seg, rows * w[:, None])
        if combiner == "sum":
            return _t(num)
        den = torch.zeros(n, dtype=params.dtype).index_add_(0, seg, w)
        if combiner == "sqrtn":
            den = torch.zeros(n, dtype=params.dtype).index_add_(0, seg, w * w).sqrt()
        return _t(num / den[:, None])

    @staticmethod
    def moments(x, axes, keep_dims=False, keepdims=False):
        x = _t(x)
        kd = keep_dims or keepdims
        mean = x.mean(dim=axes, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=axes, keepdim=kd)
        return _t(mean if kd else mean.squeeze(axes)), _t(var)

    @staticmethod
    def dropout(x, keep_prob=1.0, **_kw):
        if keep_prob >= 1.0:
            return _t(x)
        return _t(torch.nn.functional.dropout(_t(x), 1.0 - keep_prob, True))

    @staticmethod
    def l2_loss(x):
        return _t((_t(x) ** 2).sum() / 2)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(logits=None, labels=None):
        z, y = _t(logits), _t(labels)
        return _t(torch.clamp(z, min=0) - z * y + torch.log1p(torch.exp(-z.abs())))

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(labels=None, logits=None):
        lp = torch.log_softmax(_t(logits), dim=-1)
        return _t(-lp.gather(-1, _t(labels).long().unsqueeze(-1)).squeeze(-1))


nn = _NN()


class _Sparse(object):
    @staticmethod
    def to_dense(sp, default_value=0):
        out = torch.full(sp.dense_shape, default_value, dtype=sp.values.dtype)
        if sp.values.numel():
            out[tuple(sp.indices.t().long())] = sp.values
        return _t(out)


sparse = _Sparse()
sparse_tensor_to_dense = _Sparse.to_dense


class _Layers(object):
    @staticmethod
    def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
              name=None, reuse=None, **_kw):
        """tf.layers.dense: variable scope `name` or the uniquified default 'dense'; kernel
        glorot-uniform [in, units], bias zeros; applied to the last axis."""
        inputs = _t(inputs)
        with variable_scope(name, default_name="dense", reuse=reuse):
            k = get_variable("kernel", [inputs.shape[-1], int(units)],
                             initializer=kernel_initializer or glorot_uniform_initializer())
            y = torch.matmul(inputs, k)
            if use_bias:
                y = y + get_variable("bias", [int(units)], initializer=bias_initializer or zeros_initializer())
        y = _t(y)
        return activation(y) if activation is not None else y

    @staticmethod
    def dropout(inputs, rate=0.5, training=False, name=None, **_kw):
        if not training or rate <= 0:
            return _t(inputs)
        return _t(torch.nn.functional.dropout(_t(inputs), rate, True))


layers = _Layers()


class _ContribLayers(object):
    @staticmethod
    def xavier_initializer(uniform=True, **_kw):
        return glorot_uniform_initializer()

    @staticmethod
    def l2_regularizer(scale):
        return lambda v: scale * (v ** 2).sum() / 2


class _Contrib(object):
    layers = _ContribLayers()


contrib = _Contrib()


class _KerasBackend(object):
    @staticmethod
    def epsilon():
        return 1e-7

    @staticmethod
    def sparse_categorical_crossentropy(target=None, output=None, from_logits=False, axis=-1):
        """tensorflow/python/keras/backend.py (1.12): clip to [eps, 1-eps], log, then
        sparse_softmax_cross_entropy_with_logits(labels=int64(target), logits=that)."""
        output = _t(output)
        if not from_logits:
            output = torch.log(torch.clamp(output, 1e-7, 1 - 1e-7))
        labels = _t(target).to(torch.int64).reshape(-1)
        return nn.sparse_softmax_cross_entropy_with_logits(labels=labels, logits=output)


class _Keras(object):
    backend = _KerasBackend()


keras = _Keras()


class _Losses(object):
    @staticmethod
    def get_regularization_losses():
        return list(_S.reg_losses)


losses = _Losses()


class _GraphKeys(object):
    UPDATE_OPS = "update_ops"
    LOCAL_VARIABLES = "local_variables"


GraphKeys = _GraphKeys()


def get_collection(*a, **k):
    return []


def add_to_collection(*a, **k):
    return None


class _Train(object):
    """Optimizer factories are only constructed, never stepped, by the code the shim runs."""
    class _Opt(object):
        def __init__(self, learning_rate=None, *a, **k):
            self.learning_rate = learning_rate

    AdamOptimizer = GradientDescentOptimizer = AdadeltaOptimizer = AdagradOptimizer = _Opt
    FtrlOptimizer = RMSPropOptimizer = _Opt


train = _Train()
