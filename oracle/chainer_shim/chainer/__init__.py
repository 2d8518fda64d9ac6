"""Stand-in `chainer` package — TEST INFRASTRUCTURE ONLY.

Chainer cannot be installed in the build container (no network, no wheel), and
all of the reference's conv arithmetic lives inside it.  This
package implements exactly the handful of Chainer symbols that the reference's
hot-path modules touch, in plain numpy float32, so that the reference `.py`
files under /root/reference can be imported and executed UNMODIFIED to produce
golden vectors (oracle/gen_golden.py).  It is never imported by the product
package `iago_b200`.

Semantics restated (Chainer v4+ CPU path):
  * Convolution2D = cross-correlation, NCHW / OIHW, zero padding, optional bias
  * Bias(shape)   = broadcast add along axis 1
  * Linear        = x.reshape(N, -1) @ W.T (+ b)
  * softmax       = exp(x - max) / sum along `axis`
  * dropout       = identity when chainer.config.train is False
  * load_npz      = '/'-separated keys -> child links by attribute name

Parity status at this boundary: UNPINNED (the reference has no tests or golden
vectors and does not pin a Chainer version) — see DESIGN.md.
"""
import contextlib

import numpy as np

from . import functions  # noqa: F401
from . import links  # noqa: F401
from . import serializers  # noqa: F401
from . import cuda  # noqa: F401
from . import optimizers  # noqa: F401
from . import optimizer_hooks  # noqa: F401
from .core import Chain, Link, Variable  # noqa: F401


class _Config:
    train = True
    enable_backprop = True
    dtype = np.float32


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)
