"""A minimal stand-in for the TensorFlow-1 graph API, just large enough to EXECUTE the reference's own
agents/utils.py, agents/policies.py and agents/models.py (fc, lstm, *Policy._build_net, prepare_loss,
IA2C / MA2C / IQL forward / add_transition / backward) in a container without TensorFlow.

TEST INFRASTRUCTURE ONLY (used by tests/golden/gen_learner_golden.py to produce committed fixtures).

How it works: every `tf.*` call builds a lazy node; `Session.run(fetches, feed_dict)` evaluates the nodes with
float64 torch tensors, so the graph STRUCTURE (which op on which operand, in which order, with which variable
names) is entirely the reference's code, while the arithmetic of each primitive is torch's:
  matmul, + - * /, slicing, relu, sigmoid, tanh, softmax, log, clip_by_value, one_hot, reduce_sum/mean/max,
  square, split, concat, squeeze, expand_dims, where, stop_gradient.
`tf.gradients` differentiates the reference-built loss with torch.autograd.  Three pieces are TF *semantics*
restated here from TensorFlow 1.12's documented kernels (they are not in the reference repo):
  clip_by_global_norm :  g * clip / max(global_norm, clip)
  RMSPropOptimizer    :  ms <- ms + (g^2 - ms)(1 - decay), ms starts at ONE; mom <- momentum*mom + lr*g/sqrt(ms + eps);
                         var <- var - mom          (training_ops ApplyRMSProp; epsilon inside the sqrt)
  AdamOptimizer       :  lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m, v moments; var <- var - lr_t*m/(sqrt(v)+1e-8)
`trainable_variables(scope)` filters by re.match on the variable name, as TF does.
`train.Saver.save` writes {variable name: array} to `<path>-<step>.npz` (the name list of a TF checkpoint).
"""
from __future__ import annotations

import builtins
import contextlib
import re
import types

import numpy as np
import torch

DT = torch.float64


class Dimension:
    def __init__(self, v):
        self.value = v

    def __floordiv__(self, o):
        return self.value // int(o)

    def __mul__(self, o):
        return self.value * int(o)

    __rmul__ = __mul__

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __repr__(self):
        return "Dimension(%r)" % (self.value,)


class TShape:
    def __init__(self, dims):
        self.dims = [Dimension(d) for d in dims]

    def __getitem__(self, i):
        return self.dims[i]

    def __len__(self):
        return len(self.dims)

    def as_list(self):
        return [d.value for d in self.dims]


class _Graph:
    def __init__(self):
        self.variables = {}          # full name -> Variable (creation order preserved)
        self.scopes = []
        self.optimizers = []


_G = _Graph()


def _val(x, ctx):
    if isinstance(x, Tensor):
        return x.eval(ctx)
    if isinstance(x, (list, tuple)):
        return [_val(e, ctx) for e in x]
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x.astype(np.float64) if x.dtype.kind == "f" else x)
    return x


def _example(x):
    if isinstance(x, Tensor):
        return x.example
    if isinstance(x, (list, tuple)):
        return [_example(e) for e in x]
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x.astype(np.float64) if x.dtype.kind == "f" else x)
    return x


class Tensor:
    """Lazy node.  `fn(*evaluated_inputs)` -> torch tensor."""

    def __init__(self, fn, inputs, name="op"):
        self.fn, self.inputs, self.name = fn, inputs, name
        try:                 # static shape inference on example operands (None dims stand in as 2)
            with torch.no_grad():
                ex = fn(*[_example(a) for a in inputs])
            self.example = ex.detach() if isinstance(ex, torch.Tensor) else torch.as_tensor(ex)
        except Exception:    # unknown batch dims that do not line up statically: shape stays unknown
            self.example = None

    @property
    def shape(self):
        return TShape(list(self.example.shape))

    def get_shape(self):
        return self.shape

    def eval(self, ctx):
        k = id(self)
        if k not in ctx:
            ctx[k] = self.fn(*[_val(a, ctx) for a in self.inputs])
        return ctx[k]

    # operators ---------------------------------------------------------------------------------
    def __add__(self, o): return Tensor(lambda a, b: a + b, [self, o], "add")
    def __radd__(self, o): return Tensor(lambda a, b: b + a, [self, o], "add")
    def __sub__(self, o): return Tensor(lambda a, b: a - b, [self, o], "sub")
    def __rsub__(self, o): return Tensor(lambda a, b: b - a, [self, o], "sub")
    def __mul__(self, o): return Tensor(lambda a, b: a * b, [self, o], "mul")
    def __rmul__(self, o): return Tensor(lambda a, b: b * a, [self, o], "mul")
    def __truediv__(self, o): return Tensor(lambda a, b: a / b, [self, o], "div")
    def __neg__(self): return Tensor(lambda a: -a, [self], "neg")
    def __getitem__(self, idx): return Tensor(lambda a: a[idx], [self], "slice")
    __hash__ = object.__hash__


class Placeholder(Tensor):
    def __init__(self, dtype, shape):
        self.dtype = dtype
        shp = [2 if d is None else int(d) for d in (shape if shape is not None else [])]
        tdt = {"float32": DT, "int32": torch.int64, "bool": torch.bool}[dtype]
        self.fn, self.inputs, self.name = None, [], "placeholder"
        self.example = torch.zeros(shp, dtype=tdt)

    def eval(self, ctx):
        if id(self) not in ctx:
            raise KeyError("placeholder not fed")
        return ctx[id(self)]


class Variable(Tensor):
    def __init__(self, name, value):
        self.name = name
        self.fn, self.inputs = None, []
        self.tensor = torch.tensor(np.asarray(value, dtype=np.float64), dtype=DT, requires_grad=True)
        self.example = self.tensor.detach()

    def eval(self, ctx):
        return self.tensor

    def numpy(self):
        return self.tensor.detach().numpy().copy()

    def assign(self, value):
        with torch.no_grad():
            self.tensor.copy_(torch.as_tensor(np.asarray(value, dtype=np.float64)))


# ---- graph / scopes / variables -----------------------------------------------------------------
def reset_default_graph():
    global _G
    _G = _Graph()


def set_random_seed(seed):
    pass          # no TF-side randomness is used by the reference graphs (initialisers draw from numpy's global RNG)


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _G.scopes.append(name)
    try:
        yield
    finally:
        _G.scopes.pop()


def get_variable(name, shape=None, initializer=None, dtype=None):
    full = "/".join(_G.scopes + [name])
    if full in _G.variables:
        return _G.variables[full]
    shape = [int(s) for s in shape]
    v = Variable(full, initializer(shape, np.float32))
    _G.variables[full] = v
    return v


def constant_initializer(c):
    return lambda shape, dtype, partition_info=None: np.full(shape, c, np.float32)


def trainable_variables(scope=None):
    vs = list(_G.variables.values())
    if scope is None:
        return vs
    return [v for v in vs if re.match(scope, v.name)]


def global_variables_initializer():
    return Tensor(lambda: torch.zeros(()), [], "init")


def placeholder(dtype, shape=None):
    return Placeholder(dtype, shape)


# ---- ops ----------------------------------------------------------------------------------------
def matmul(a, b): return Tensor(lambda x, y: x @ y, [a, b], "matmul")
def tanh(x): return Tensor(torch.tanh, [x], "tanh")
def log(x): return Tensor(torch.log, [x], "log")
def square(x): return Tensor(lambda a: a * a, [x], "square")
def clip_by_value(x, lo, hi): return Tensor(lambda a: torch.clamp(a, lo, hi), [x], "clip")
def squeeze(x): return Tensor(lambda a: a.squeeze(), [x], "squeeze")
def expand_dims(x, axis): return Tensor(lambda a: a.unsqueeze(axis), [x], "expand_dims")
def stop_gradient(x): return Tensor(lambda a: a.detach(), [x], "stop_gradient")
def where(c, a, b): return Tensor(lambda cc, x, y: torch.where(cc, x, y), [c, a, b], "where")
def one_hot(idx, depth): return Tensor(lambda i: torch.nn.functional.one_hot(i.long(), int(depth)).to(DT), [idx], "one_hot")


def _reduce(f, name):
    def op(x, axis=None):
        if axis is None:
            return Tensor(lambda a: f(a), [x], name)
        return Tensor(lambda a: f(a, dim=axis), [x], name)
    return op


reduce_sum = _reduce(torch.sum, "reduce_sum")
reduce_mean = _reduce(torch.mean, "reduce_mean")


def reduce_max(x, axis=None):
    if axis is None:
        return Tensor(lambda a: a.max(), [x], "reduce_max")
    return Tensor(lambda a: a.max(dim=axis).values, [x], "reduce_max")


def split(axis=0, num_or_size_splits=1, value=None):
    n = int(num_or_size_splits)
    size = value.example.shape[axis] // n
    return [Tensor(lambda a, i=i: a.narrow(axis, i * size, size), [value], "split") for i in range(n)]


def concat(values=None, axis=0, name=None):
    vals = list(values)
    return Tensor(lambda *xs: torch.cat(list(xs), dim=axis), vals, "concat")


def gradients(ys, xs):
    group = {}

    def all_grads(ctx):
        k = ("grads", id(ys), tuple(id(x) for x in xs))
        if k not in ctx:
            loss = ys.eval(ctx)
            gs = torch.autograd.grad(loss, [x.tensor for x in xs], allow_unused=True, retain_graph=True)
            ctx[k] = [torch.zeros_like(x.tensor) if g is None else g.detach() for g, x in zip(gs, xs)]
        return ctx[k]

    class GradTensor(Tensor):
        def __init__(self, i):
            self.i, self.fn, self.inputs, self.name = i, None, [], "gradient"
            self.example = xs[i].example

        def eval(self, ctx):
            return all_grads(ctx)[self.i]
    return [GradTensor(i) for i in range(len(xs))]


def clip_by_global_norm(t_list, clip_norm):
    norm = Tensor(lambda *gs: torch.sqrt(sum((g * g).sum() for g in gs)), list(t_list), "global_norm")
    clipped = [Tensor(lambda g, n: g * (clip_norm / torch.clamp(n, min=clip_norm)), [g, norm], "clip") for g in t_list]
    return clipped, norm


class _TrainOp(Tensor):
    def __init__(self, apply_fn, deps):
        self.apply_fn, self.deps = apply_fn, deps
        self.fn, self.inputs, self.name = None, [], "train"
        self.example = torch.zeros(())

    def eval(self, ctx):           # evaluated LAST by Session.run
        if id(self) not in ctx:
            vals = [_val(d, ctx) for d in self.deps]
            self.apply_fn(vals)
            ctx[id(self)] = torch.zeros(())
        return ctx[id(self)]


class RMSPropOptimizer:
    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
        self.lr, self.decay, self.momentum, self.eps = learning_rate, decay, momentum, epsilon
        self.slots = {}
        _G.optimizers.append(self)

    def apply_gradients(self, grads_and_vars):
        gv = list(grads_and_vars)
        for _, v in gv:
            self.slots[v.name] = [torch.ones_like(v.tensor), torch.zeros_like(v.tensor)]   # rms = 1, momentum = 0

        def apply(vals):
            lr, grads = vals[0], vals[1:]
            lr = float(lr)
            with torch.no_grad():
                for g, (_, v) in zip(grads, gv):
                    ms, mom = self.slots[v.name]
                    ms += (g * g - ms) * (1.0 - self.decay)
                    mom.mul_(self.momentum).add_(lr * g / torch.sqrt(ms + self.eps))
                    v.tensor -= mom
        return _TrainOp(apply, [self.lr] + [g for g, _ in gv])


class AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.slots, self.t = {}, 0

    def apply_gradients(self, grads_and_vars):
        gv = list(grads_and_vars)
        for _, v in gv:
            self.slots[v.name] = [torch.zeros_like(v.tensor), torch.zeros_like(v.tensor)]

        def apply(vals):
            lr, grads = float(vals[0]), vals[1:]
            self.t += 1
            lr_t = lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
            with torch.no_grad():
                for g, (_, v) in zip(grads, gv):
                    m, s = self.slots[v.name]
                    m += (g - m) * (1.0 - self.b1)
                    s += (g * g - s) * (1.0 - self.b2)
                    v.tensor -= lr_t * m / (torch.sqrt(s) + self.eps)
        return _TrainOp(apply, [self.lr] + [g for g, _ in gv])


class Saver:
    def __init__(self, max_to_keep=5):
        self.vars = list(_G.variables.values())      # variables existing NOW (as tf.train.Saver())

    def save(self, sess, path, global_step=None):
        out = "%s-%d.npz" % (path, int(global_step))
        np.savez(out, **{v.name: v.numpy().astype(np.float32) for v in self.vars})
        return out

    def restore(self, sess, path):
        z = np.load(path + ".npz")
        for v in self.vars:
            v.assign(z[v.name])


class Session:
    def __init__(self, config=None):
        pass

    def run(self, fetches, feed_dict=None):
        ctx = {}
        for ph, val in (feed_dict or {}).items():
            a = np.asarray(val)
            if ph.dtype == "float32":
                t = torch.as_tensor(a.astype(np.float64))
            elif ph.dtype == "int32":
                t = torch.as_tensor(a.astype(np.int64))
            else:
                t = torch.as_tensor(a.astype(np.bool_))
            ctx[id(ph)] = t
        flat = []

        def collect(f):
            if isinstance(f, (list, tuple)):
                for e in f:
                    collect(e)
            else:
                flat.append(f)
        collect(fetches)
        for f in flat:                      # values first, side effects last
            if not isinstance(f, _TrainOp):
                f.eval(ctx)
        for f in flat:
            if isinstance(f, _TrainOp):
                f.eval(ctx)

        def out(f):
            if isinstance(f, (list, tuple)):
                return [out(e) for e in f]
            if isinstance(f, _TrainOp):
                return None
            v = f.eval(ctx)
            return v.detach().numpy().copy() if isinstance(v, torch.Tensor) else v
        return out(fetches)


def make_module():
    """The object to install as sys.modules['tensorflow']."""
    tf = types.ModuleType("tensorflow")
    g = globals()
    for k in ("reset_default_graph", "set_random_seed", "variable_scope", "get_variable", "constant_initializer",
              "trainable_variables", "global_variables_initializer", "placeholder", "matmul", "tanh", "log", "square",
              "clip_by_value", "squeeze", "expand_dims", "stop_gradient", "where", "one_hot", "reduce_sum",
              "reduce_mean", "reduce_max", "split", "concat", "gradients", "clip_by_global_norm", "Session"):
        setattr(tf, k, g[k])
    tf.float32, tf.int32, tf.bool = "float32", "int32", "bool"
    tf.__version__ = "1.12-shim"
    tf.ConfigProto = lambda **k: None
    nn = types.SimpleNamespace(
        relu=lambda x: Tensor(torch.relu, [x], "relu"),
        sigmoid=lambda x: Tensor(torch.sigmoid, [x], "sigmoid"),
        softmax=lambda x: Tensor(lambda a: torch.softmax(a, dim=-1), [x], "softmax"),
        conv1d=None, conv2d=None)
    tf.nn = nn
    tf.train = types.SimpleNamespace(RMSPropOptimizer=RMSPropOptimizer, AdamOptimizer=AdamOptimizer, Saver=Saver)
    tf.summary = types.SimpleNamespace(scalar=lambda name, t: Tensor(lambda: torch.zeros(()), [], "summary"),
                                       merge=lambda ls: Tensor(lambda: torch.zeros(()), [], "summary"))
    tf.shim_graph = lambda: _G
    return tf


_ = builtins  # (kept: `bool` is never shadowed at module level, tf.bool is set on the module object only)
