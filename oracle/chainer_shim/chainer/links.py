"""chainer.links stand-ins: Convolution2D, Bias, Linear (fp32 numpy)."""
import numpy as np

from .core import Link, Variable, as_array


def _im2col(x, k, pad):
    n, c, h, w = x.shape
    xp = np.zeros((n, c, h + 2 * pad, w + 2 * pad), dtype=x.dtype)
    xp[:, :, pad:pad + h, pad:pad + w] = x
    oh, ow = h + 2 * pad - k + 1, w + 2 * pad - k + 1
    col = np.empty((n, c, k, k, oh, ow), dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            col[:, :, i, j] = xp[:, :, i:i + oh, j:j + ow]
    return col


class Convolution2D(Link):
    """Cross-correlation, NCHW/OIHW, zero padding (Chainer CPU: im2col + tensordot)."""

    def __init__(self, in_channels, out_channels, ksize=None, stride=1, pad=0, nobias=False):
        super().__init__()
        self.out_channels, self.ksize, self.pad, self.nobias = out_channels, ksize, pad, nobias
        with self.init_scope():
            self.W = None
            if not nobias:
                self.b = None

    def __call__(self, x):
        x = as_array(x)
        col = _im2col(x, self.ksize, self.pad)
        y = np.tensordot(col, self.W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)
        if not self.nobias:
            y += self.b
        return Variable(np.ascontiguousarray(np.rollaxis(y, 3, 1)))


class Bias(Link):
    def __init__(self, axis=1, shape=None):
        super().__init__()
        self.axis = axis
        with self.init_scope():
            self.b = None

    def __call__(self, x):
        x = as_array(x)
        shp = [1] * x.ndim
        shp[self.axis:self.axis + self.b.ndim] = self.b.shape
        return Variable(x + self.b.reshape(shp))


class Linear(Link):
    def __init__(self, in_size, out_size=None, nobias=False):
        super().__init__()
        self.nobias = nobias
        with self.init_scope():
            self.W = None
            if not nobias:
                self.b = None

    def __call__(self, x):
        x = as_array(x)
        y = x.reshape(len(x), -1).dot(self.W.T).astype(x.dtype, copy=False)
        if not self.nobias:
            y += self.b
        return Variable(y)


class Classifier(Link):
    def __init__(self, predictor, *a, **k):
        super().__init__()
        with self.init_scope():
            self.predictor = predictor
