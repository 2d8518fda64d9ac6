"""Variable / Link / Chain stand-ins (see package docstring)."""
import contextlib

import numpy as np


class Variable:
    def __init__(self, data=None):
        if isinstance(data, Variable):
            data = data.data
        self.data = data

    @property
    def array(self):
        return self.data

    @property
    def shape(self):
        return self.data.shape

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Variable(self.data.reshape(shape))

    def __len__(self):
        return len(self.data)


def as_array(x):
    return x.data if isinstance(x, Variable) else np.asarray(x)


class Link:
    def __init__(self):
        object.__setattr__(self, "_params", [])
        object.__setattr__(self, "_children", [])
        object.__setattr__(self, "_in_scope", False)

    @contextlib.contextmanager
    def init_scope(self):
        object.__setattr__(self, "_in_scope", True)
        try:
            yield
        finally:
            object.__setattr__(self, "_in_scope", False)

    def __setattr__(self, name, value):
        if getattr(self, "_in_scope", False):
            if isinstance(value, Link):
                self._children.append(name)
            elif value is None or isinstance(value, np.ndarray):
                if name not in self._params:
                    self._params.append(name)
        object.__setattr__(self, name, value)

    def add_param_name(self, name):
        if name not in self._params:
            self._params.append(name)

    def namedparams(self, prefix=""):
        for p in self._params:
            yield prefix + "/" + p, self, p
        for c in self._children:
            yield from getattr(self, c).namedparams(prefix + "/" + c)

    def to_gpu(self, *a, **k):  # pragma: no cover - never used on the hot path
        raise RuntimeError("chainer stand-in has no GPU path")

    def to_cpu(self):
        return self

    def cleargrads(self):
        pass


class Chain(Link):
    pass
