"""tf18shim: an EAGER stand-in for the subset of TensorFlow 1.8's Python API that nabu's hot path uses.

TEST INFRASTRUCTURE ONLY (tests/golden/make_tf18shim_golden.py imports it; nothing under nabu_b200/ does).

Why it exists: the reference (vrenkens/nabu) is Python-2 + TensorFlow-1.8 code and neither is in the build image, so
the reference's own modules could not be executed and the oracle was "parity unpinned".  With this package first on
sys.path as `tensorflow`, the reference's UNMODIFIED sources (loaded from /root/reference by tests/golden/py2ref.py)
run: models/model.py, ed_encoders/{listener,dblstm}.py, ed_decoders/{speller,rnn_decoder,dnn_decoder}.py,
components/{layer,ops,attention,rnn_cell,beam_search_decoder}.py, trainers/loss_functions.py,
decoders/{ctc_decoder,beam_search_decoder}.py.  Everything those files author themselves (layer stacking, pyramid
stacking, the location-aware / windowed attention score, probability functions, the projection wrapper, the target
shifting, the loss normalisation, the whole beam search) is therefore the REFERENCE'S code; what this package restates
is TensorFlow's side of the calls, op by op, from the published r1.8 sources (file and symbol cited at each op):
tensors are torch tensors (fp64 by default, fp32 with set_float_bits(32)), so gradients come from torch autograd.

What is NOT pinned by this: TensorFlow's own kernels (the restatement here is a second, independent one next to
oracle/nabu_oracle.py - agreement of two restatements, not a run of TF), and `tf.nn.ctc_beam_search_decoder`,
which is delegated to the oracle (only the reference's wrapper around it is exercised).

Graph-mode behaviour that matters and is emulated: variable-scope naming (default-name uniquification through the
per-store scope counts, reset when a scope closes - tensorflow/python/ops/variable_scope.py), AUTO_REUSE,
`tf.layers` objects capturing their scope at the first call, and loop bodies being traced ONCE (the counters are
restored at the start of every eager iteration so that an iteration re-creates the same names).
"""
import builtins
import collections
import contextlib
import re

import numpy as np
import torch

_FLOAT = [torch.float64]


def set_float_bits(bits):
    """working precision of every float tensor: 64 (truth for logits / loss / gradients) or 32 (TF's own arithmetic,
    used for the beam-search ids)"""
    _FLOAT[0] = {64: torch.float64, 32: torch.float32}[bits]


# ---------------------------------------------------------------------------------------------------- dtypes, shapes
class DType(object):
    def __init__(self, name, kind):
        self.name, self.kind = name, kind

    @property
    def max(self):
        return {'float32': float(np.finfo(np.float32).max), 'float64': float(np.finfo(np.float64).max),
                'int32': 2 ** 31 - 1, 'int64': 2 ** 63 - 1}[self.name]

    @property
    def min(self):
        return -self.max if self.kind == 'f' else -self.max - 1

    @property
    def torch(self):
        return {'f': _FLOAT[0], 'b': torch.bool}.get(self.kind) or {'int32': torch.int32, 'int64': torch.int64}[self.name]

    def __repr__(self):
        return 'tf.' + self.name


float32, float64 = DType('float32', 'f'), DType('float64', 'f')
int32, int64, bool = DType('int32', 'i'), DType('int64', 'i'), DType('bool', 'b')      # noqa: A001


def _dtype_of(t):
    if t.dtype.is_floating_point:
        return float32
    return {torch.int32: int32, torch.int64: int64, torch.bool: bool}[t.dtype]


class Dimension(object):
    def __init__(self, value):
        self.value = value

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, other):
        return self.value == (other.value if isinstance(other, Dimension) else other)

    def __hash__(self):
        return hash(self.value)

    def __mul__(self, other):
        return int(self) * int(other)

    __rmul__ = __mul__

    def __add__(self, other):
        return int(self) + int(other)

    __radd__ = __add__

    def __repr__(self):
        return 'Dimension(%r)' % self.value


class TensorShape(object):
    def __init__(self, dims):
        self._dims = [d if isinstance(d, Dimension) else Dimension(d) for d in (dims or [])]

    @property
    def ndims(self):
        return len(self._dims)

    def as_list(self):
        return [d.value for d in self._dims]

    def __len__(self):
        return len(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __getitem__(self, key):
        return TensorShape(self._dims[key]) if isinstance(key, slice) else self._dims[key]

    def concatenate(self, other):
        return TensorShape(self._dims + list(TensorShape(list(other))._dims if not isinstance(other, TensorShape)
                                             else other._dims))

    def __repr__(self):
        return 'TensorShape(%r)' % self.as_list()


# ------------------------------------------------------------------------------------------------------------ tensors
def _t(x, dtype=None):
    """anything -> torch tensor (python floats in the working precision, python ints int32 like tf.constant)"""
    if isinstance(x, Tensor):
        t = x.t
    elif isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, Dimension):
        t = torch.tensor(int(x), dtype=torch.int32)
    elif isinstance(x, (builtins.bool, np.bool_)):
        t = torch.tensor(builtins.bool(x))
    elif isinstance(x, (int, np.integer)):
        t = torch.tensor(int(x), dtype=torch.int32)
    elif isinstance(x, (float, np.floating)):
        t = torch.tensor(float(x), dtype=_FLOAT[0])
    elif isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
    elif isinstance(x, (list, tuple)) or type(x).__name__ in ('dict_values', 'dict_keys'):
        x = list(x)
        if not x:
            t = torch.zeros(0, dtype=_FLOAT[0])
        else:
            parts = [_t(v) for v in x]
            kind = torch.result_type(parts[0], parts[0])
            for p in parts[1:]:
                kind = torch.promote_types(kind, p.dtype)
            t = torch.stack([p.to(kind) for p in parts])
    else:
        raise TypeError('tf18shim: cannot convert %r' % type(x))
    if t.dtype.is_floating_point and t.dtype != _FLOAT[0]:
        t = t.to(_FLOAT[0])
    if dtype is not None:
        t = t.to(dtype.torch if isinstance(dtype, DType) else dtype)
    return t


def _ints(x):
    """a shape / multiples / permutation argument -> list of python ints"""
    if isinstance(x, Tensor):
        return [int(v) for v in x.t.reshape(-1).tolist()]
    if isinstance(x, TensorShape):
        return x.as_list()
    if isinstance(x, (int, np.integer, Dimension)):
        return [int(x)]
    return [int(v) for v in x]


def _pair(a, b):
    a, b = _t(a), _t(b)
    if a.dtype != b.dtype:
        kind = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(kind), b.to(kind)
    return a, b


class _Op(object):
    def __init__(self, name):
        self.name = name


class Tensor(object):
    __array_priority__ = 100

    def __init__(self, t, name=None):
        self.t = t
        self._name = name

    # --- static information
    @property
    def shape(self):
        return TensorShape(list(self.t.shape))

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        pass

    @property
    def dtype(self):
        return _dtype_of(self.t)

    @property
    def name(self):
        return (self._name or 'tensor') + ':0'

    @property
    def op(self):
        return _Op(self._name or 'tensor')

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __int__(self):
        return int(self.t)

    __index__ = __int__

    def __float__(self):
        return float(self.t)

    def __bool__(self):
        return builtins.bool(self.t)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return (Tensor(v) for v in self.t)

    def __repr__(self):
        return 'tf18shim.Tensor(%r)' % (self.t,)

    __hash__ = object.__hash__

    # --- operators (python-2 semantics: `/` on integers floors, as tf.div does)
    def __add__(self, o):
        return Tensor(torch.add(*_pair(self, o)))

    def __radd__(self, o):
        return Tensor(torch.add(*_pair(o, self)))

    def __sub__(self, o):
        return Tensor(torch.sub(*_pair(self, o)))

    def __rsub__(self, o):
        return Tensor(torch.sub(*_pair(o, self)))

    def __mul__(self, o):
        return Tensor(torch.mul(*_pair(self, o)))

    def __rmul__(self, o):
        return Tensor(torch.mul(*_pair(o, self)))

    def __truediv__(self, o):
        return div(self, o)

    def __rtruediv__(self, o):
        return div(o, self)

    __div__, __rdiv__ = __truediv__, __rtruediv__

    def __floordiv__(self, o):
        return floor_div(self, o)

    def __mod__(self, o):
        return mod(self, o)

    def __pow__(self, o):
        return Tensor(torch.pow(*_pair(self, o)))

    def __rpow__(self, o):
        return Tensor(torch.pow(*_pair(o, self)))

    def __neg__(self):
        return Tensor(-self.t)

    def __lt__(self, o):
        return Tensor(torch.lt(*_pair(self, o)))

    def __le__(self, o):
        return Tensor(torch.le(*_pair(self, o)))

    def __gt__(self, o):
        return Tensor(torch.gt(*_pair(self, o)))

    def __ge__(self, o):
        return Tensor(torch.ge(*_pair(self, o)))

    def __getitem__(self, key):
        def conv(k):
            if isinstance(k, (Tensor, Dimension)):
                return int(k)
            if isinstance(k, slice):
                return slice(*[None if v is None else int(v) for v in (k.start, k.stop, k.step)])
            return k
        key = tuple(conv(k) for k in key) if isinstance(key, tuple) else conv(key)
        return Tensor(self.t[key])


class Variable(Tensor):
    def __init__(self, t, name, trainable=True):
        Tensor.__init__(self, t, name)
        self.trainable = trainable

    @property
    def initializer(self):
        return no_op()

    def assign(self, value):
        with torch.no_grad():
            self.t.copy_(_t(value))
        return self


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = convert_to_tensor(indices), convert_to_tensor(values), \
            convert_to_tensor(dense_shape)


SparseTensorValue = collections.namedtuple('SparseTensorValue', ('indices', 'values', 'dense_shape'))


def convert_to_tensor(x, dtype=None, name=None):
    return x if isinstance(x, Tensor) and dtype is None else Tensor(_t(x, dtype))


constant = convert_to_tensor
newaxis = None


# ------------------------------------------------------------------------------------------ variable scopes, variables
AUTO_REUSE = 'AUTO_REUSE'


class GraphKeys(object):
    GLOBAL_VARIABLES = 'variables'
    TRAINABLE_VARIABLES = 'trainable_variables'


class VariableScope(object):
    """tensorflow/python/ops/variable_scope.py: VariableScope (the fields the reference reads)"""

    def __init__(self, reuse, name='', initializer=None, constraint=None, custom_getter=None, **_):
        self.reuse, self.name, self.initializer, self.constraint, self.custom_getter = \
            reuse, name, initializer, constraint, custom_getter

    @property
    def original_name_scope(self):
        return self.name + '/'


class _Store(object):
    def __init__(self):
        self.vars = collections.OrderedDict()
        self.counts = {}
        self.stack = [VariableScope(None, '')]
        self.rng = np.random.RandomState(0)


_S = _Store()


def reset_default_graph():
    """a new graph: names start over; the variables stay (they play the role of the checkpoint that every graph of
    the reference restores)"""
    _S.counts = {}
    _S.stack = [VariableScope(None, '')]


def reset_all(seed=0):
    _S.vars.clear()
    reset_default_graph()
    _S.rng = np.random.RandomState(seed)


def set_random_seed(seed):
    _S.rng = np.random.RandomState(seed)


def get_variable_scope():
    return _S.stack[-1]


def _unique_scope(prefix):
    """variable_scope.py: _get_unique_variable_scope"""
    cur = get_variable_scope().name
    name = cur + '/' + prefix if cur else prefix
    if _S.counts.get(name, 0) == 0:
        return prefix
    idx = 1
    while _S.counts.get('%s_%d' % (name, idx), 0) > 0:
        idx += 1
    return '%s_%d' % (prefix, idx)


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, values=None, reuse=None, custom_getter=None, initializer=None,
                   **_):
    """variable_scope.py: variable_scope / _pure_variable_scope.  A VariableScope object is entered by its own full
    name (not nested under the current one) and the scope counts are put back when it is left; a string is nested,
    counted, and the counts of its sub-scopes are zeroed when it is left."""
    cur = get_variable_scope()
    if name_or_scope is None:
        name_or_scope = _unique_scope(default_name)
    if isinstance(name_or_scope, VariableScope):
        old = dict(_S.counts)
        full = name_or_scope.name
        _S.counts[full] = _S.counts.get(full, 0) + 1
        new = VariableScope(reuse if reuse is not None else name_or_scope.reuse, full, name_or_scope.initializer,
                            name_or_scope.constraint, custom_getter or name_or_scope.custom_getter)
        _S.stack.append(new)
        try:
            yield new
        finally:
            _S.stack.pop()
            _S.counts = old
    else:
        full = cur.name + '/' + name_or_scope if cur.name else name_or_scope
        _S.counts[full] = _S.counts.get(full, 0) + 1
        new = VariableScope(reuse if reuse is not None else cur.reuse, full, initializer or cur.initializer,
                            cur.constraint, custom_getter or cur.custom_getter)
        _S.stack.append(new)
        try:
            yield new
        finally:
            _S.stack.pop()
            for k in _S.counts:
                if k.startswith(full + '/'):
                    _S.counts[k] = 0


@contextlib.contextmanager
def _traced_once():
    """a graph-mode loop body is traced once: every eager iteration starts from the same scope counts"""
    snapshot = dict(_S.counts)

    def again():
        _S.counts = dict(snapshot)
    try:
        yield again
    finally:
        pass


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
    yield name or default_name or ''


@contextlib.contextmanager
def control_dependencies(_):
    yield


def zeros_initializer(dtype=None):
    return 'zeros'


def ones_initializer(dtype=None):
    return 'ones'


def constant_initializer(value=0.0, dtype=None):
    return ('constant', value)


def glorot_uniform_initializer(seed=None, dtype=None):
    return 'glorot_uniform'


def _initial_value(shape, initializer):
    shape = _ints(shape)
    if initializer is None or initializer == 'glorot_uniform':
        # tensorflow/python/ops/init_ops.py: VarianceScaling(1.0, fan_avg, uniform) - get_variable's default
        if len(shape) < 1:
            fan_in = fan_out = 1
        elif len(shape) == 1:
            fan_in = fan_out = shape[0]
        else:
            rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
        limit = np.sqrt(6.0 / (fan_in + fan_out))
        return _S.rng.uniform(-limit, limit, size=shape)
    if initializer == 'zeros':
        return np.zeros(shape)
    if initializer == 'ones':
        return np.ones(shape)
    if isinstance(initializer, tuple) and initializer[0] == 'constant':
        return np.full(shape, initializer[1])
    raise NotImplementedError('tf18shim: initializer %r' % (initializer,))


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, constraint=None, **_):
    """variable_scope.py: get_variable.  Every scope of the reference is AUTO_REUSE, so an existing name is returned
    and a new one is created; values: the named initializer drawn from the store's seeded generator, fp32-rounded (the
    values travel as a float32 checkpoint)."""
    scope = get_variable_scope()
    full = scope.name + '/' + name if scope.name else name
    if full in _S.vars:
        return _S.vars[full]
    if scope.reuse is True:
        raise ValueError('tf18shim: variable %s does not exist (reuse=True)' % full)
    if isinstance(initializer, Tensor):
        value = initializer.t.detach().double().numpy()
    else:
        value = _initial_value(shape if shape is not None else [], initializer or scope.initializer)
    kind = dtype.torch if isinstance(dtype, DType) else _FLOAT[0]
    if kind.is_floating_point:
        t = torch.tensor(np.asarray(value, np.float32).astype(np.float64), dtype=kind, requires_grad=trainable)
    else:
        t = torch.tensor(np.asarray(value), dtype=kind)
    var = Variable(t, full, trainable)
    _S.vars[full] = var
    return var


def get_collection(key, scope=None):
    """ops.py: Graph.get_collection - re.match(scope, name) over the collection, creation order"""
    out = []
    for name, var in _S.vars.items():
        if key == GraphKeys.TRAINABLE_VARIABLES and not var.trainable:
            continue
        if scope is None or re.match(scope, name + ':0'):
            out.append(var)
    return out


def global_variables():
    return list(_S.vars.values())


def trainable_variables():
    return [v for v in _S.vars.values() if v.trainable]


def cast_variables_(bits):
    """re-type the stored variables (switching the working precision between two graphs)"""
    set_float_bits(bits)
    for var in _S.vars.values():
        if var.t.dtype.is_floating_point:
            var.t = var.t.detach().to(_FLOAT[0]).requires_grad_(var.trainable)


# ----------------------------------------------------------------------------------------------------------- basic ops
def shape(x, name=None, out_type=None):
    return Tensor(torch.tensor(list(_t(x).shape), dtype=torch.int32))


def size(x, name=None):
    return Tensor(torch.tensor(_t(x).numel(), dtype=torch.int32))


def rank(x):
    return Tensor(torch.tensor(_t(x).dim(), dtype=torch.int32))


def cast(x, dtype, name=None):
    if isinstance(x, SparseTensor):                               # math_ops.cast: the values of a SparseTensor
        return SparseTensor(x.indices, cast(x.values, dtype), x.dense_shape)
    return Tensor(_t(x).to(dtype.torch))


def to_float(x, name=None):
    return cast(x, float32)


def to_int32(x, name=None):
    return cast(x, int32)


def identity(x, name=None):
    return Tensor(_t(x))


def no_op(name=None):
    return None


def zeros(shape, dtype=float32, name=None):
    return Tensor(torch.zeros(_ints(shape), dtype=dtype.torch))


def ones(shape, dtype=float32, name=None):
    return Tensor(torch.ones(_ints(shape), dtype=dtype.torch))


def fill(dims, value, name=None):
    v = _t(value)
    return Tensor(torch.full(_ints(dims), v.item(), dtype=v.dtype))


def zeros_like(x, dtype=None, name=None, optimize=True):
    t = _t(x)
    return Tensor(torch.zeros_like(t, dtype=dtype.torch if dtype is not None else t.dtype))


def ones_like(x, dtype=None, name=None, optimize=True):
    t = _t(x)
    return Tensor(torch.ones_like(t, dtype=dtype.torch if dtype is not None else t.dtype))


def range(start, limit=None, delta=1, dtype=None, name=None):        # noqa: A001
    if limit is None:
        start, limit = 0, start
    return Tensor(torch.arange(int(start), int(limit), int(delta), dtype=(dtype or int32).torch))


def concat(values, axis, name=None):
    parts = [_t(v) for v in values]
    parts = [p.reshape(1) if p.dim() == 0 else p for p in parts]
    kind = parts[0].dtype
    for p in parts[1:]:
        kind = torch.promote_types(kind, p.dtype)
    return Tensor(torch.cat([p.to(kind) for p in parts], int(axis)))


def stack(values, axis=0, name=None):
    parts = [_t(v) for v in values]
    kind = parts[0].dtype
    for p in parts[1:]:
        kind = torch.promote_types(kind, p.dtype)
    return Tensor(torch.stack([p.to(kind) for p in parts], int(axis)))


def unstack(value, num=None, axis=0, name=None):
    return [Tensor(v) for v in torch.unbind(_t(value), int(axis))]


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    t = _t(value)
    if isinstance(num_or_size_splits, (int, np.integer)):
        return [Tensor(v) for v in torch.chunk(t, int(num_or_size_splits), int(axis))]
    return [Tensor(v) for v in torch.split(t, _ints(num_or_size_splits), int(axis))]


def transpose(a, perm=None, name=None):
    t = _t(a)
    return Tensor(t.permute(*(_ints(perm) if perm is not None else reversed(builtins.range(t.dim())))))


def reshape(tensor, shape, name=None):
    return Tensor(_t(tensor).reshape(_ints(shape)))


def expand_dims(input, axis=None, name=None, dim=None):        # noqa: A002
    return Tensor(_t(input).unsqueeze(int(axis if axis is not None else dim)))


def squeeze(input, axis=None, name=None, squeeze_dims=None):        # noqa: A002
    axis = axis if axis is not None else squeeze_dims
    t = _t(input)
    if axis is None:
        return Tensor(t.squeeze())
    for a in sorted(_ints(axis), reverse=True):
        t = t.squeeze(a)
    return Tensor(t)


def tile(input, multiples, name=None):        # noqa: A002
    return Tensor(_t(input).repeat(*_ints(multiples)))


def gather(params, indices, validate_indices=None, name=None, axis=0):
    p, i = _t(params), _t(indices).long()
    return Tensor(torch.index_select(p, int(axis), i.reshape(-1)).reshape(
        list(p.shape[:int(axis)]) + list(i.shape) + list(p.shape[int(axis) + 1:])))


def gather_nd(params, indices, name=None):
    p, i = _t(params), _t(indices).long()
    k = i.shape[-1]
    return Tensor(p[tuple(i[..., d] for d in builtins.range(k))])


def where(condition, x=None, y=None, name=None):
    c = _t(condition)
    if x is None:
        return Tensor(torch.nonzero(c).to(torch.int64))          # row-major order, [n, rank] int64 like TF
    a, b = _pair(x, y)
    if c.dim() == 1 and a.dim() > 1:                             # tf.where: a vector condition selects rows
        c = c.reshape([-1] + [1] * (a.dim() - 1))
    return Tensor(torch.where(c, a, b))


def pad(tensor, paddings, mode='CONSTANT', name=None, constant_values=0):
    t = _t(tensor)
    for axis, (before, after) in enumerate(_t(paddings).tolist()):
        parts = [t]
        for n, front in ((before, True), (after, False)):
            if n > 0:
                shp = list(t.shape)
                shp[axis] = int(n)
                block = torch.full(shp, constant_values, dtype=t.dtype)
                parts = [block] + parts if front else parts + [block]
        if len(parts) > 1:
            t = torch.cat(parts, axis)
    return Tensor(t)


def one_hot(indices, depth, on_value=None, off_value=None, axis=None, dtype=None, name=None):
    i = _t(indices).long()
    hot = i.unsqueeze(-1) == torch.arange(int(depth))
    return Tensor(hot.to((dtype or float32).torch))


def sequence_mask(lengths, maxlen=None, dtype=bool, name=None):
    n = _t(lengths).long()
    m = int(maxlen) if maxlen is not None else int(n.max())
    return Tensor((torch.arange(m) < n.unsqueeze(-1)).to(dtype.torch))


def reverse_sequence(input, seq_lengths, seq_axis=None, batch_axis=None, name=None, seq_dim=None, batch_dim=None):  # noqa
    """array_ops.reverse_sequence, batch axis 0, sequence axis 1"""
    t, n = _t(input), _t(seq_lengths).long()
    assert (seq_axis if seq_axis is not None else seq_dim) == 1 and (batch_axis or batch_dim or 0) == 0
    T = t.shape[1]
    pos = torch.arange(T).unsqueeze(0).expand(t.shape[0], T)
    idx = torch.where(pos < n.unsqueeze(1), n.unsqueeze(1) - 1 - pos, pos)
    return Tensor(torch.gather(t, 1, idx.reshape(list(idx.shape) + [1] * (t.dim() - 2)).expand_as(t)))


def cumsum(x, axis=0, exclusive=False, reverse=False, name=None):
    assert not exclusive and not reverse
    return Tensor(torch.cumsum(_t(x), int(axis)))


def _reduce(fn):
    def op(input_tensor, axis=None, keepdims=None, name=None, reduction_indices=None, keep_dims=None):
        t = _t(input_tensor)
        axis = axis if axis is not None else reduction_indices
        keep = builtins.bool(keepdims or keep_dims)
        if axis is None:
            out = fn(t, list(builtins.range(t.dim())), keep) if t.dim() else t
        else:
            out = fn(t, _ints(axis), keep)
        return Tensor(out)
    return op


reduce_sum = _reduce(lambda t, a, k: torch.sum(t, a, keepdim=k))
reduce_mean = _reduce(lambda t, a, k: torch.mean(t, a, keepdim=k))
reduce_max = _reduce(lambda t, a, k: torch.amax(t, a, keepdim=k))
reduce_min = _reduce(lambda t, a, k: torch.amin(t, a, keepdim=k))
reduce_all = _reduce(lambda t, a, k: torch.all(t.reshape(-1)) if not k else None)
reduce_any = _reduce(lambda t, a, k: torch.any(t.reshape(-1)) if not k else None)


def _unary(fn):
    return lambda x, name=None: Tensor(fn(_t(x)))


def _binary(fn):
    return lambda x, y, name=None: Tensor(fn(*_pair(x, y)))


tanh, sigmoid, square, sqrt, exp, log, ceil, floor = (_unary(f) for f in (
    torch.tanh, torch.sigmoid, torch.square, torch.sqrt, torch.exp, torch.log, torch.ceil, torch.floor))
rsqrt, logical_not = _unary(torch.rsqrt), _unary(torch.logical_not)
abs, negative = _unary(torch.abs), _unary(torch.neg)        # noqa: A001
add, subtract, multiply, maximum, minimum = (_binary(f) for f in (
    torch.add, torch.sub, torch.mul, torch.maximum, torch.minimum))
less, less_equal, greater, greater_equal, equal, not_equal = (_binary(f) for f in (
    torch.lt, torch.le, torch.gt, torch.ge, torch.eq, torch.ne))
logical_and, logical_or, logical_xor = (_binary(f) for f in (torch.logical_and, torch.logical_or, torch.logical_xor))


def div(x, y, name=None):
    """math_ops.div: python-2 division - true division of floats, floor division of integers"""
    a, b = _pair(x, y)
    return Tensor(a / b if a.dtype.is_floating_point else torch.div(a, b, rounding_mode='floor'))


def floor_div(x, y, name=None):
    a, b = _pair(x, y)
    return Tensor(torch.div(a, b, rounding_mode='floor'))


floordiv = floor_div


def mod(x, y, name=None):
    return Tensor(torch.remainder(*_pair(x, y)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _pair(a, b)
    return Tensor(torch.matmul(a.transpose(-1, -2) if transpose_a else a, b.transpose(-1, -2) if transpose_b else b))


def tensordot(a, b, axes, name=None):
    a, b = _pair(a, b)
    return Tensor(torch.tensordot(a, b, axes))


def norm(tensor, ord='euclidean', axis=None, keepdims=None, name=None, keep_dims=None):        # noqa: A002
    """linalg_ops.norm, euclidean"""
    assert ord in ('euclidean', 2)
    t = _t(tensor)
    keep = builtins.bool(keepdims or keep_dims)
    return Tensor(torch.sqrt(torch.sum(t * t, dim=_ints(axis), keepdim=keep)) if axis is not None else torch.sqrt(torch.sum(t * t)))


def clip_by_value(t, lo, hi, name=None):
    return Tensor(torch.clamp(_t(t), float(lo), float(hi)))


def assert_less(x, y, message=None, **_):
    assert builtins.bool(torch.all(torch.lt(*_pair(x, y)))), message


def assert_less_equal(x, y, message=None, **_):
    assert builtins.bool(torch.all(torch.le(*_pair(x, y)))), message


def sparse_tensor_to_dense(sp, default_value=0, validate_indices=True, name=None):
    out = torch.full(_ints(sp.dense_shape), default_value, dtype=sp.values.t.dtype)
    idx = sp.indices.t.long()
    out[tuple(idx[:, d] for d in builtins.range(idx.shape[1]))] = sp.values.t
    return Tensor(out)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    return Tensor(torch.from_numpy(_S.rng.normal(mean, stddev, size=_ints(shape))).to(_FLOAT[0]))


def edit_distance(hypothesis, truth, normalize=True, name='edit_distance'):
    raise NotImplementedError('tf18shim: edit_distance (evaluation bookkeeping, outside the pinned path)')


# ------------------------------------------------------------------------------------------- TensorArray, while_loop
class TensorArray(object):
    """tensor_array_ops.TensorArray, eager: a python list (write returns self, as the flow-carrying object does)"""

    def __init__(self, dtype, size=0, dynamic_size=None, element_shape=None, infer_shape=True, name=None, **_):
        self.dtype = dtype
        self._items = [None] * int(size)

    def write(self, index, value, name=None):
        i = int(index)
        new = TensorArray(self.dtype)
        new._items = list(self._items) + [None] * (i + 1 - len(self._items))
        new._items[i] = convert_to_tensor(value)
        return new

    def read(self, index, name=None):
        return self._items[int(index)]

    def size(self, name=None):
        return Tensor(torch.tensor(len(self._items), dtype=torch.int32))

    def stack(self, name=None):
        return stack(self._items, 0)

    def unstack(self, value, name=None):
        new = TensorArray(self.dtype)
        new._items = unstack(value, axis=0)
        return new

    def split(self, value, lengths, name=None):
        new = TensorArray(self.dtype)
        new._items = [Tensor(v) for v in torch.split(_t(value), _ints(lengths), 0)]
        return new


def while_loop(cond, body, loop_vars, shape_invariants=None, parallel_iterations=10, back_prop=True,
               swap_memory=False, name=None, maximum_iterations=None):
    loop_vars = list(loop_vars)
    with _traced_once() as again:
        n = 0
        while builtins.bool(_t(cond(*loop_vars))) and (maximum_iterations is None or n < int(maximum_iterations)):
            again()
            out = body(*loop_vars)
            loop_vars = list(out) if isinstance(out, (list, tuple)) else [out]
            n += 1
    return loop_vars


def cond(pred, true_fn=None, false_fn=None, name=None, **_):
    return true_fn() if builtins.bool(_t(pred)) else false_fn()


# -------------------------------------------------------------------------------------------------------- sub-modules
from . import nn, layers, contrib, summary, train        # noqa: E402,F401
