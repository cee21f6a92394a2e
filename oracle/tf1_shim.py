"""TEST INFRASTRUCTURE: a minimal TensorFlow-1 graph interpreter on torch float64, so that the reference's own
graph-definition code (modules/base/{losses,initializers,optimizers,mapping}.py, models/*, approaches/*: the
`_define_variables` / `_define_embed_graph` methods and `session.run([loss, optimizer], feed_dict)`) can be EXECUTED,
unmodified, in a container without TensorFlow.  `scripts/make_golden_path_i.py` installs this module as `tensorflow`,
imports the reference from /root/reference/src and writes tests/golden/path_i_*.npz; nothing under openea_b200/
imports it.

What comes from the reference when a golden is generated: which variables exist and how they are initialised /
normalised, every lookup, the whole loss expression, which optimiser instance minimises which loss.
What is restated HERE (from the TF 1.x documentation / op definitions, not executable offline — the residual
"unpinned" part of path (i)'s parity):
  * op semantics: reduce_sum / reduce_mean, pow, square, abs (gradient sign(x), 0 at 0), relu (gradient 0 at 0),
    maximum (ties send the gradient to the first argument), log, exp, sigmoid, softplus, matmul(transpose_b),
    nn.l2_normalize(x, axis, epsilon=1e-12) = x · rsqrt(max(Σ_axis x², epsilon)) — over ALL elements when axis is None;
  * gradients of a loss w.r.t. a variable are dense (an embedding_lookup of l2_normalize(var) is dense in TF too; for
    lookups straight into a variable TF's sparse Adagrad / SGD / Adam updates equal the dense rule; only sparse
    Adadelta differs, which no shipped configuration uses);
  * optimiser rules: GradientDescent; Adagrad (initial_accumulator_value 0.1); Adam (β₁ .9, β₂ .999, ε 1e-8,
    lr_t = lr·√(1−β₂ᵗ)/(1−β₁ᵗ), ε outside the root); Adadelta (ρ .95, ε 1e-8).
Arithmetic is float64 so that a golden is the mathematical value of the reference's graph to ~1e-15.
"""
import builtins
import contextlib
import math
import types

import numpy as np
import torch

DT = torch.float64
builtins_slice = builtins.slice
int32, int64, float32, float64 = "int32", "int64", "float32", "float64"
_VARIABLES = []


def reset_default_graph():
    _VARIABLES.clear()
    _NAMED.clear()
    _LAYER_COUNT.clear()


def _wrap(x):
    if isinstance(x, Tensor):
        return x
    return Tensor(lambda: _const(x), (), "const")


def _const(x):
    if isinstance(x, torch.Tensor) or type(x).__name__ == "SparseValue":
        return x
    a = np.asarray(x)
    return torch.as_tensor(a, dtype=DT if a.dtype.kind == "f" else None)


class Tensor:
    """A node of the lazy graph: fn(*evaluated inputs) → torch tensor."""

    def __init__(self, fn, inputs, name="op"):
        self.fn, self.inputs, self.name = fn, tuple(inputs), name

    def _eval(self, env):
        key = id(self)
        if key not in env:
            env[key] = self.fn(*[i._eval(env) for i in self.inputs])
        return env[key]

    def eval(self, session=None, feed_dict=None):
        return Session().run(self, feed_dict)

    def _bin(self, other, f, name, swap=False):
        a, b = (_wrap(other), self) if swap else (self, _wrap(other))
        return Tensor(f, (a, b), name)

    def __getitem__(self, idx):
        return Tensor(lambda a: a[idx], (self,), "getitem")

    @property
    def shape(self):
        """Static shape, by evaluating the node once without feeds (only for nodes that do not depend on a placeholder)."""
        return tuple(self._eval({}).shape)

    def __add__(self, o): return self._bin(o, torch.add, "add")
    def __radd__(self, o): return self._bin(o, torch.add, "add", True)
    def __sub__(self, o): return self._bin(o, torch.sub, "sub")
    def __rsub__(self, o): return self._bin(o, torch.sub, "sub", True)
    def __mul__(self, o): return self._bin(o, _mul, "mul")
    def __rmul__(self, o): return self._bin(o, _mul, "mul", True)
    def __truediv__(self, o): return self._bin(o, torch.div, "div")
    def __rtruediv__(self, o): return self._bin(o, torch.div, "div", True)
    def __neg__(self): return Tensor(torch.neg, (self,), "neg")
    def __pow__(self, p): return pow(self, p)


def _mul(a, b):
    """Dense product, or SparseTensor * dense with TF's broadcasting of a [n, 1] / [1, n] dense operand over the
    stored entries (alinet.py:668-669)."""
    if type(a).__name__ != "SparseValue":
        return torch.mul(a, b)
    row, col = a.indices[:, 0], a.indices[:, 1]
    if b.dim() == 2 and b.shape[1] == 1:
        scale = b[row, 0]
    elif b.dim() == 2 and b.shape[0] == 1:
        scale = b[0, col]
    else:
        scale = b if b.dim() == 0 else b[row, col]
    return SparseValue(a.indices, a.values * scale, a.shape)


class Placeholder(Tensor):
    def __init__(self, dtype, shape=None, name=None):
        super().__init__(None, (), name or "placeholder")
        self.dtype = dtype

    def _eval(self, env):
        if id(self) not in env:
            raise KeyError("placeholder %s was not fed" % self.name)
        return env[id(self)]


class Variable(Tensor):
    def __init__(self, initial_value, name=None, dtype=None, trainable=True):
        super().__init__(None, (), name or "Variable")
        if isinstance(initial_value, Tensor):              # tf.Variable(tf.truncated_normal(...)): evaluate the initialiser
            initial_value = initial_value._eval({}).detach().numpy()
        self.value = torch.tensor(np.asarray(initial_value, dtype=np.float64), dtype=DT, requires_grad=True)
        self.slots = {}
        taken = {v.name for v in _VARIABLES}
        base, n = self.name, 0
        while self.name in taken:                          # TF uniquifies: Variable, Variable_1, …
            n += 1
            self.name = "%s_%d" % (base, n)
        if trainable:
            _VARIABLES.append(self)

    def _eval(self, env):
        return self.value

    def assign_numpy(self, array):
        self.value = torch.tensor(np.asarray(array, dtype=np.float64), dtype=DT, requires_grad=True)
        self.slots = {}


_NAMED = {}           # "name:0" → placeholder (tf.get_default_graph().get_tensor_by_name, string keys of a feed_dict)


def placeholder(dtype, shape=None, name=None):
    ph = Placeholder(dtype, shape, name)
    if name is not None:
        _NAMED[name + ":0"] = ph
    return ph


class _DefaultPlaceholder(Placeholder):
    def __init__(self, default, shape=None, name=None):
        super().__init__(None, shape, name)
        self.default = default

    def _eval(self, env):
        return env[id(self)] if id(self) in env else _const(self.default)


def placeholder_with_default(input, shape=None, name=None):       # noqa: A002
    return _DefaultPlaceholder(input, shape, name)


class SparseValue:
    """What a sparse placeholder is fed with / a tf.SparseTensor holds: COO indices [nnz, 2], values, dense shape."""

    def __init__(self, indices, values, dense_shape):
        self.indices = torch.as_tensor(np.asarray(indices)).long().reshape(-1, 2)
        self.values = values if isinstance(values, torch.Tensor) else torch.as_tensor(np.asarray(values), dtype=DT)
        self.shape = tuple(int(x) for x in dense_shape)


def sparse_placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name or "sparse_placeholder")


def sparse_tensor_dense_matmul(sp_a, b, name=None):
    def f(a, dense):
        out = torch.zeros(a.shape[0], dense.shape[1], dtype=DT)
        return out.index_add(0, a.indices[:, 0], a.values[:, None] * dense[a.indices[:, 1]])
    return Tensor(f, (_wrap(sp_a), _wrap(b)), "sparse_tensor_dense_matmul")


def add_n(inputs, name=None):
    out = _wrap(inputs[0])
    for x in inputs[1:]:
        out = out + x
    return out


def _random(kind):
    def op(shape, mean=0.0, stddev=1.0, minval=0.0, maxval=1.0, dtype=None, seed=None, name=None):
        if kind == "normal":
            return _wrap(_truncated_normal(stddev=stddev, mean=mean)(list(shape)))
        return _wrap(_random_uniform(minval=minval, maxval=maxval)(list(shape)))
    return op


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    return _wrap(_truncated_normal(stddev=stddev, mean=mean)(list(shape)))


def random_uniform(shape, minval=0.0, maxval=None, dtype=None, seed=None, name=None):
    return _wrap(_random_uniform(minval=minval, maxval=maxval)(list(shape)))


def zeros(shape, dtype=None, name=None): return _wrap(np.zeros(shape, dtype=np.float64))
def ones(shape, dtype=None, name=None): return _wrap(np.ones(shape, dtype=np.float64))


class GraphKeys:
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"


def get_collection(key, scope=None):
    return list(_VARIABLES)


class _Graph:
    def get_tensor_by_name(self, name):
        return _NAMED[name]                  # KeyError when absent, like TF


def get_default_graph():
    return _Graph()


class SparseTensor(Tensor):
    """tf.SparseTensor(indices, values, dense_shape): every component may be a graph node; .indices / .values /
    .dense_shape give back what was passed (the reference uses r_mat.values as lookup ids, rdgcn.py:205)."""

    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape

        def f(i, v, shape):
            return SparseValue(i, v if v.dtype.is_floating_point else v.to(DT), [int(x) for x in shape])
        super().__init__(f, (_wrap(indices), _wrap(values), _wrap(dense_shape)), "SparseTensor")


class _SparseResult(Tensor):
    """A sparse-valued node whose components can be taken apart again (weights.indices, weights.values, …)."""
    indices = property(lambda self: Tensor(lambda a: a.indices, (self,), "indices"))
    values = property(lambda self: Tensor(lambda a: a.values, (self,), "values"))
    dense_shape = property(lambda self: Tensor(lambda a: torch.as_tensor(a.shape), (self,), "dense_shape"))


def sparse_add(a, b, name=None):
    def f(x, y):
        assert torch.equal(x.indices, y.indices), "sparse_add is only needed for operands with one pattern"
        return SparseValue(x.indices, x.values + y.values, x.shape)
    return _SparseResult(f, (_wrap(a), _wrap(b)), "sparse_add")


def sparse_reshape(sp_input, shape, name=None):
    return _wrap(sp_input)


def unstack(value, num=None, axis=0, name=None):
    x = _wrap(value)
    n = num if num is not None else x.shape[axis]
    return [Tensor(lambda a, i=i: a.select(axis, i), (x,), "unstack") for i in range(n)]


def stack(values, axis=0, name=None):
    return Tensor(lambda *v: torch.stack(v, dim=axis), tuple(_wrap(v) for v in values), "stack")


def reverse(tensor, axis, name=None):
    return Tensor(lambda a: torch.flip(a, dims=list(axis)), (_wrap(tensor),), "reverse")


def _static(x):
    """A Python int from a constant node (loop counters, slice offsets): evaluated without feeds."""
    return int(x._eval({})) if isinstance(x, Tensor) else int(x)


def slice(input_, begin, size, name=None):                       # noqa: A001
    b, n = [_static(v) for v in begin], [_static(v) for v in size]

    def f(a):
        idx = tuple(builtins_slice(lo, None if cnt == -1 else lo + cnt) for lo, cnt in zip(b, n))
        return a[idx]
    return Tensor(f, (_wrap(input_),), "slice")


def greater(x, y, name=None):
    return Tensor(lambda a, b: a > b, (_wrap(x), _wrap(y)), "greater")


def while_loop(cond, body, loop_vars, **kw):
    """Unrolled at graph-construction time: the reference's only loop (attre.py:89-107) counts a constant down, so its
    condition can be evaluated without feeds while the body keeps building lazy nodes."""
    variables = list(loop_vars)
    while bool(cond(*variables)._eval({})):
        variables = list(body(*variables))
    return variables


def tile(x, multiples, name=None):
    return Tensor(lambda a: a.repeat(*multiples), (_wrap(x),), "tile")


def sparse_softmax(sp_input, name=None):
    """Softmax over the stored entries of every row (entries, not coordinates: duplicates of one (i, j) are separate
    entries).  ASSUMPTION: rows are taken whole — TF documents the op for canonically ordered input; the reference
    passes triples in list order (rdgcn.py:20-42)."""
    def f(a):
        row = a.indices[:, 0]
        mx = torch.full((a.shape[0],), -float("inf"), dtype=DT).scatter_reduce(0, row, a.values, "amax")
        ex = torch.exp(a.values - mx[row])
        den = torch.zeros(a.shape[0], dtype=DT).index_add(0, row, ex)
        return SparseValue(a.indices, ex / den[row], a.shape)
    return Tensor(f, (_wrap(sp_input),), "sparse_softmax")


def expand_dims(x, axis, name=None):
    return Tensor(lambda a: a.unsqueeze(axis), (_wrap(x),), "expand_dims")


def transpose(x, perm=None, name=None):
    return Tensor(lambda a: a.t() if perm is None else a.permute(*perm), (_wrap(x),), "transpose")


def concat(values, axis, name=None):
    return Tensor(lambda *v: torch.cat(v, dim=axis), tuple(_wrap(v) for v in values), "concat")


_LAYER_COUNT = {}


def _conv1d(inputs, filters, kernel_size, use_bias=True, **kw):
    """tf.layers.conv1d with kernel_size 1 on [1, n, c]: a dense map with its own glorot-uniform kernel [1, c, filters]
    and zero bias, variables named conv1d[_i]/kernel, conv1d[_i]/bias in creation order.  The channel count is read by
    evaluating the input once at graph-construction time (it must not depend on a placeholder)."""
    assert kernel_size == 1
    x = _wrap(inputs)
    c = int(x._eval({}).shape[-1])
    idx = _LAYER_COUNT.get("conv1d", 0)
    _LAYER_COUNT["conv1d"] = idx + 1
    scope = "conv1d" if idx == 0 else "conv1d_%d" % idx
    kernel = Variable(_xavier(uniform=True)([c, filters]).reshape(1, c, filters), name=scope + "/kernel")
    if not use_bias:
        return Tensor(lambda a, k: a @ k[0], (x, kernel), scope)
    bias = Variable(np.zeros(filters), name=scope + "/bias")
    return Tensor(lambda a, k, b: a @ k[0] + b, (x, kernel, bias), scope)


layers = types.SimpleNamespace(conv1d=_conv1d)
class _BatchNormalization:
    """tf.keras.layers.BatchNormalization() called in a TF-1 graph without `training`: inference mode — the moving
    statistics (mean 0, variance 1 at initialisation, never updated by the reference's train op) normalise, so
    y = γ·x / √(1 + ε) + β with ε = 1e-3; γ, β are created at the first call and shared by later calls."""

    def __init__(self, epsilon=1e-3, **kw):
        self.epsilon, self.gamma, self.beta = epsilon, None, None
        idx = _LAYER_COUNT.get("bn", 0)
        _LAYER_COUNT["bn"] = idx + 1
        self.scope = "batch_normalization" if idx == 0 else "batch_normalization_%d" % idx

    def __call__(self, inputs, training=None):
        x = _wrap(inputs)
        if self.gamma is None:
            c = x.shape[-1]
            self.gamma = Variable(np.ones(c), name=self.scope + "/gamma")
            self.beta = Variable(np.zeros(c), name=self.scope + "/beta")
        eps = self.epsilon
        return Tensor(lambda a, g, b: a * (g / math.sqrt(1.0 + eps)) + b, (x, self.gamma, self.beta), self.scope)


keras = types.SimpleNamespace(layers=types.SimpleNamespace(BatchNormalization=_BatchNormalization),
                              activations=types.SimpleNamespace(
    get=lambda name: {"relu": nn.relu, "tanh": tanh, None: None}[name], relu=lambda x: nn.relu(x), tanh=lambda x: tanh(x)))

summary = types.SimpleNamespace(histogram=lambda *a, **kw: None, scalar=lambda *a, **kw: None)


def constant(value, dtype=None, name=None, shape=None):
    return _wrap(value)


def get_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True, **kw):
    # a regularizer only registers a loss in a collection; none of the reference's losses reads that collection
    return Variable(initializer(list(shape)), name=name, dtype=dtype, trainable=trainable)


def trainable_variables():
    return list(_VARIABLES)


@contextlib.contextmanager
def name_scope(name):
    yield


variable_scope = name_scope


def _unary(f, name):
    return lambda x, name_=None, **kw: Tensor(f, (_wrap(x),), name)


abs = _unary(torch.abs, "abs")               # noqa: A001  (TF's names)
square = _unary(torch.square, "square")
log = _unary(torch.log, "log")
exp = _unary(torch.exp, "exp")
sigmoid = _unary(torch.sigmoid, "sigmoid")
tanh = _unary(torch.tanh, "tanh")


def pow(x, y, name=None):                    # noqa: A001
    return Tensor(lambda a, b: torch.pow(a, b), (_wrap(x), _wrap(y)), "pow")


def add(x, y, name=None): return _wrap(x) + y
def subtract(x, y, name=None): return _wrap(x) - y
def multiply(x, y, name=None): return _wrap(x) * y


def maximum(x, y, name=None):
    return Tensor(lambda a, b: torch.where(a >= b, a + 0 * b, b + 0 * a), (_wrap(x), _wrap(y)), "maximum")


def cast(x, dtype=None, name=None):
    def f(a):
        if type(a).__name__ == "SparseValue":
            return a
        return a.to(DT) if dtype in (float32, float64) else a.long()
    return Tensor(f, (_wrap(x),), "cast")


def _reduce(f):
    def op(x, axis=None, keepdims=False, keep_dims=False, name=None):
        keep = bool(keepdims or keep_dims)
        return Tensor(lambda a: f(a) if axis is None else f(a, dim=axis, keepdim=keep), (_wrap(x),), f.__name__)
    return op


reduce_sum = _reduce(torch.sum)
reduce_mean = _reduce(torch.mean)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None, **kw):
    def f(x, y):
        return (x.t() if transpose_a else x) @ (y.t() if transpose_b else y)
    return Tensor(f, (_wrap(a), _wrap(b)), "matmul")


def reshape(x, shape, name=None):
    return Tensor(lambda a: a.reshape(shape), (_wrap(x),), "reshape")


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    axis = dim if axis is None else axis

    def f(a):
        ss = (a * a).sum() if axis is None else (a * a).sum(dim=axis, keepdim=True)
        return a * torch.rsqrt(torch.clamp(ss, min=epsilon))
    return Tensor(f, (_wrap(x),), "l2_normalize")


def _embedding_lookup(params, ids, name=None):
    return Tensor(lambda p, i: p[torch.as_tensor(i).long()], (_wrap(params), _wrap(ids)), "embedding_lookup")


def _dropout(x, keep_prob=None, rate=None, noise_shape=None, seed=None, name=None):
    """Every shipped configuration trains with dropout 0: only the identity is supported (anything else is random)."""
    def f(a, kp):
        assert float(kp) == 1.0, "dropout with keep_prob != 1 is random and cannot be put in a golden"
        return a
    return Tensor(f, (_wrap(x), _wrap(1.0 - rate if keep_prob is None else keep_prob)), "dropout")


def _leaky_relu(x, alpha=0.2, name=None):
    return Tensor(lambda a: torch.where(a > 0, a, alpha * a), (_wrap(x),), "leaky_relu")


nn = types.SimpleNamespace(
    dropout=_dropout, leaky_relu=_leaky_relu, bias_add=lambda x, b, name=None: _wrap(x) + b,
    softmax=lambda x, axis=-1, name=None: Tensor(lambda a: torch.softmax(a, dim=axis), (_wrap(x),), "softmax"),
    embedding_lookup=_embedding_lookup, l2_normalize=_l2_normalize,
    relu=_unary(torch.relu, "relu"), softplus=_unary(torch.nn.functional.softplus, "softplus"),
    sigmoid=sigmoid, tanh=tanh)


# ---- initialisers (the goldens overwrite the values; shapes and dtypes are what matters) ------------------------
def _gen():
    g = torch.Generator()
    g.manual_seed(1234)
    return g


def _truncated_normal(stddev=1.0, mean=0.0, **kw):
    def init(shape):
        x = torch.empty(*shape, dtype=DT)
        torch.nn.init.trunc_normal_(x, mean=mean, std=stddev, a=mean - 2 * stddev, b=mean + 2 * stddev, generator=_gen())
        return x.numpy()
    return init


def _random_uniform(minval=0, maxval=None, **kw):
    hi = 1.0 if maxval is None else maxval
    return lambda shape: (torch.rand(*shape, dtype=DT, generator=_gen()) * (hi - minval) + minval).numpy()


def _orthogonal(gain=1.0, **kw):
    def init(shape):
        x = torch.empty(*shape, dtype=DT)
        torch.nn.init.orthogonal_(x, gain=gain, generator=_gen())
        return x.numpy()
    return init


def _xavier(uniform=True, **kw):
    def init(shape):
        fan = (shape[0] + shape[1]) / 2.0
        if uniform:
            lim = math.sqrt(3.0 / fan)
            return ((torch.rand(*shape, dtype=DT, generator=_gen()) * 2 - 1) * lim).numpy()
        return _truncated_normal(stddev=math.sqrt(1.3 / fan))(shape)
    return init


glorot_uniform_initializer = lambda **kw: _xavier(uniform=True)
zeros_initializer = lambda **kw: (lambda shape: np.zeros(shape))
ones_initializer = lambda **kw: (lambda shape: np.ones(shape))
initializers = types.SimpleNamespace(truncated_normal=_truncated_normal, random_uniform=_random_uniform,
                                     orthogonal=_orthogonal)
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=_xavier,
                                                             l2_regularizer=lambda scale=0.0, **kw: None))
truncated_normal_initializer = _truncated_normal
random_uniform_initializer = _random_uniform


# ---- optimisers ------------------------------------------------------------------------------------------------
def _variables_of(node, seen=None, out=None):
    seen = set() if seen is None else seen
    out = [] if out is None else out
    if id(node) in seen:
        return out
    seen.add(id(node))
    if isinstance(node, Variable):
        out.append(node)
    for i in node.inputs:
        _variables_of(i, seen, out)
    return out


class _TrainOp(Tensor):
    def __init__(self, optimizer, loss, variables):
        super().__init__(None, (), "train_op")
        self.optimizer, self.loss, self.variables = optimizer, loss, variables


class _Optimizer:
    def __init__(self, learning_rate, **kw):
        self.lr = float(learning_rate)
        self.kw = kw
        self.key = id(self)          # slot variables belong to the optimiser INSTANCE
        self.t = 0

    def compute_gradients(self, loss, var_list=None):
        reach = _variables_of(loss)
        if var_list is not None:
            allowed = {id(v) for v in var_list}
            reach = [v for v in reach if id(v) in allowed]
        return (loss, reach)

    def apply_gradients(self, grads_and_vars, **kw):
        loss, variables = grads_and_vars
        return _TrainOp(self, loss, variables)

    def minimize(self, loss, var_list=None, **kw):
        return self.apply_gradients(self.compute_gradients(loss, var_list))

    def _run(self, env):
        loss = self.current.loss._eval(env)
        variables = self.current.variables
        grads = torch.autograd.grad(loss, [v.value for v in variables], allow_unused=True, retain_graph=True)
        self.t += 1
        with torch.no_grad():
            for v, g in zip(variables, grads):
                if g is not None:
                    self.update(v, g, v.slots.setdefault(self.key, {}))

    def slot(self, slots, name, like, fill=0.0):
        if name not in slots:
            slots[name] = torch.full_like(like, fill)
        return slots[name]


class GradientDescentOptimizer(_Optimizer):
    def update(self, v, g, slots):
        v.value -= self.lr * g


class AdagradOptimizer(_Optimizer):
    def update(self, v, g, slots):
        acc = self.slot(slots, "accumulator", v.value, self.kw.get("initial_accumulator_value", 0.1))
        acc += g * g
        v.value -= self.lr * g / acc.sqrt()


class AdamOptimizer(_Optimizer):
    def update(self, v, g, slots):
        b1, b2, eps = self.kw.get("beta1", 0.9), self.kw.get("beta2", 0.999), self.kw.get("epsilon", 1e-8)
        m, s = self.slot(slots, "m", v.value), self.slot(slots, "v", v.value)
        m.mul_(b1).add_(g, alpha=1 - b1)
        s.mul_(b2).addcmul_(g, g, value=1 - b2)
        lr_t = self.lr * math.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
        v.value -= lr_t * m / (s.sqrt() + eps)


class AdadeltaOptimizer(_Optimizer):
    def __init__(self, learning_rate=0.001, rho=0.95, epsilon=1e-8, **kw):
        super().__init__(learning_rate, rho=rho, epsilon=epsilon, **kw)

    def update(self, v, g, slots):
        rho, eps = self.kw["rho"], self.kw["epsilon"]
        acc, acc_up = self.slot(slots, "accum", v.value), self.slot(slots, "accum_update", v.value)
        acc.mul_(rho).addcmul_(g, g, value=1 - rho)
        up = (acc_up + eps).sqrt() / (acc + eps).sqrt() * g
        acc_up.mul_(rho).addcmul_(up, up, value=1 - rho)
        v.value -= self.lr * up


train = types.SimpleNamespace(GradientDescentOptimizer=GradientDescentOptimizer, AdagradOptimizer=AdagradOptimizer,
                              AdamOptimizer=AdamOptimizer, AdadeltaOptimizer=AdadeltaOptimizer)


# ---- session ---------------------------------------------------------------------------------------------------
class ConfigProto:
    def __init__(self, **kw):
        self.gpu_options = types.SimpleNamespace(allow_growth=False)


class _Init:
    def run(self, session=None, feed_dict=None):
        return None


def global_variables_initializer():
    return _Init()


class Session:
    def __init__(self, config=None, **kw):
        pass

    def run(self, fetches, feed_dict=None):
        env = {}
        for ph, val in (feed_dict or {}).items():
            if isinstance(ph, str):
                ph = _NAMED[ph]
            if isinstance(val, tuple) and len(val) == 3 and np.ndim(val[0]) == 2:     # (coords, values, shape)
                env[id(ph)] = SparseValue(*val)
                continue
            a = np.asarray(val)
            env[id(ph)] = torch.as_tensor(a, dtype=DT) if a.dtype.kind == "f" else torch.as_tensor(a)
        flat, rebuild = _flatten(fetches)
        values = [None] * len(flat)
        for i, f in enumerate(flat):                     # forward values first: a fetched loss is the pre-update loss
            if not isinstance(f, _TrainOp):
                out = f._eval(env)
                values[i] = out if isinstance(out, SparseValue) else out.detach().numpy().copy()
        for f in flat:
            if isinstance(f, _TrainOp):
                f.optimizer.current = f
                f.optimizer._run(env)
        for v in _VARIABLES:                              # fresh leaves for the next run's autograd graph
            v.value = v.value.detach().requires_grad_(True)
        return rebuild(values)

    def close(self):
        pass


def _flatten(fetches):
    if isinstance(fetches, dict):
        keys = list(fetches)
        return [fetches[k] for k in keys], lambda vals: dict(zip(keys, vals))
    if isinstance(fetches, (list, tuple)):
        return list(fetches), lambda vals: list(vals)
    return [fetches], lambda vals: vals[0]
