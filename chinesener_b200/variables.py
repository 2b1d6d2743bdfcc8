"""Variable store keyed by the reference's TF variable names.

The reference creates variables implicitly in TF's default graph (tf.get_variable inside
`build_graph`); checkpoints list them by scope name (serving_model/*/variables.index,
SURVEY.md §8 a9).  This store plays the same role for the plugin functions: the first
`build_graph` call creates every variable with the TF initializer it would have had, later
calls reuse them.  `state_dict()` keys are the TF names, so a converted Google-BERT /
reference checkpoint (name -> array) loads with `load_state_dict`.
"""
import contextlib
import math
from collections import OrderedDict

import torch


def truncated_normal(stddev):
    def init(shape, gen):
        t = torch.empty(shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=stddev, a=-2 * stddev, b=2 * stddev, generator=gen)
        return t
    return init


def glorot_uniform(shape, gen):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * lim


xavier = glorot_uniform  # tf.contrib.layers.xavier_initializer() is uniform by default


def zeros(shape, gen):
    return torch.zeros(shape, dtype=torch.float32)


def ones(shape, gen):
    return torch.ones(shape, dtype=torch.float32)


def constant(value):
    def init(shape, gen):
        t = torch.as_tensor(value, dtype=torch.float32).clone()
        assert tuple(t.shape) == tuple(shape)
        return t
    return init


class VariableStore:
    def __init__(self, device="cuda", seed=1234):
        self.device = torch.device(device)
        self.vars = OrderedDict()
        self.trainable = OrderedDict()
        self.gen = torch.Generator().manual_seed(seed)
        self.version = 0          # bumped whenever values change (invalidates packed-weight caches)
        self.caches = {}          # per-layer derived tensors (bf16 packs), keyed by the owning layer
        self.grads = {}           # name -> gradient tensor (views of the optimizer's flat buffer once built)
        self.global_step = 0
        self.dropout_calls = 0    # per-step counter that decorrelates the dropout layers' seeds

    def get_variable(self, name, shape, initializer, trainable=True):
        v = self.vars.get(name)
        if v is None:
            v = initializer(tuple(shape), self.gen).to(self.device).contiguous()
            self.vars[name] = v
            self.trainable[name] = trainable
            self.version += 1
        assert tuple(v.shape) == tuple(shape), f"{name}: have {tuple(v.shape)}, want {tuple(shape)}"
        return v

    def state_dict(self):
        return OrderedDict((k, v.detach().cpu()) for k, v in self.vars.items())

    def load_state_dict(self, sd, strict=False):
        for k, t in sd.items():
            t = torch.as_tensor(t, dtype=torch.float32)
            if k in self.vars:
                assert tuple(self.vars[k].shape) == tuple(t.shape), k
                self.vars[k].copy_(t.to(self.device))
            elif strict:
                raise KeyError(k)
            else:
                self.vars[k] = t.to(self.device).contiguous()
                self.trainable.setdefault(k, True)
        self.version += 1

    def touch(self):
        self.version += 1

    def grad(self, name):
        """Gradient accumulator of variable `name` (created zero on first use)."""
        g = self.grads.get(name)
        if g is None:
            g = torch.zeros_like(self.vars[name])
            self.grads[name] = g
        return g

    def trainable_names(self):
        return [n for n in self.vars if self.trainable.get(n, True)]

    def cached(self, key, builder):
        ent = self.caches.get(key)
        if ent is None or ent[0] != self.version:
            ent = (self.version, builder())
            self.caches[key] = ent
        return ent[1]

    def num_parameters(self):
        return sum(v.numel() for v in self.vars.values())


_default = None


def default_store():
    global _default
    if _default is None:
        _default = VariableStore("cuda" if torch.cuda.is_available() else "cpu")
    return _default


def set_default_store(store):
    global _default
    _default = store
    return store


@contextlib.contextmanager
def use_store(store):
    global _default
    prev = _default
    _default = store
    try:
        yield store
    finally:
        _default = prev


def get_variable(name, shape, initializer, trainable=True):
    return default_store().get_variable(name, shape, initializer, trainable)


_scope = ""


@contextlib.contextmanager
def variable_scope(name):
    """tf.variable_scope(name, reuse=tf.AUTO_REUSE) for the layer functions that take part in the multi-task plugins
    (bilstm / dense / crf_layer resolve their variable names through `scoped`)."""
    global _scope
    prev = _scope
    _scope = f"{prev}{name}/"
    try:
        yield
    finally:
        _scope = prev


def scoped(name):
    return _scope + name


class Deferred:
    """A graph tensor that is evaluated only when fetched.  The reference builds `loss` and `pred_ids` in one
    TF graph and a session run computes only what the mode fetches: PREDICT never runs the log-likelihood
    (tools/train_utils.py:181-185 exports pred_ids only).  Non-training crf_layer() returns one of these;
    `float(x)` / `x.value()` is the fetch."""

    def __init__(self, thunk):
        self._thunk, self._val = thunk, None

    def value(self):
        if self._thunk is not None:
            self._val, self._thunk = self._thunk(), None
        return self._val

    def mean(self):
        return Deferred(lambda: self.value().mean())

    def __neg__(self):
        return Deferred(lambda: -self.value())

    def __add__(self, other):
        return Deferred(lambda: self.value() + (other.value() if isinstance(other, Deferred) else other))

    __radd__ = __add__

    def __mul__(self, other):
        return Deferred(lambda: self.value() * (other.value() if isinstance(other, Deferred) else other))

    __rmul__ = __mul__

    def __float__(self):
        return float(self.value())

    def item(self):
        return float(self)
