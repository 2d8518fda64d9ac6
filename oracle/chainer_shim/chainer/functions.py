"""chainer.functions stand-ins (fp32 numpy)."""
import numpy as np

from .core import Variable, as_array


def relu(x):
    return Variable(np.maximum(as_array(x), 0))


def reshape(x, shape):
    return Variable(as_array(x).reshape(shape))


def softmax(x, axis=1):
    x = as_array(x)
    y = x - x.max(axis=axis, keepdims=True)
    np.exp(y, out=y)
    y /= y.sum(axis=axis, keepdims=True)
    return Variable(y)


def dropout(x, ratio=0.5):
    import chainer
    if chainer.config.train:
        raise RuntimeError("stand-in dropout is inference-only (chainer.config.train must be False)")
    return x if isinstance(x, Variable) else Variable(x)
