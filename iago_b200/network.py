"""SLPolicy / RolloutPolicy / Value — drop-ins for /root/reference/network.py, evaluated by the CUDA kernels.

Same call shape as the reference: model(x) with x a float32 array (N,2,8,8) of input planes
(channel 0 = opponent stones, channel 1 = mover's stones, game.py:167-174) returns
  SLPolicy / RolloutPolicy : (N,64) softmax probabilities       network.py:34-47 / 59-64
  Value                    : (N,)   scalars                      network.py:83-96
as a `Variable`-like object with `.data` (plain ndarray), so reference call sites such as
`self.model(state_var).data.reshape(64)` (mcts_self_play.py:102, MCTS.py:95) keep working.
Weights load from the reference's npz files with `load(path)` (the role of serializers.load_npz).
"""
import itertools

import numpy as np

from . import npz
from .engine import default_engine

_ids = itertools.count()
_W = (np.uint64(1) << np.arange(64, dtype=np.uint64))


class Variable:
    def __init__(self, data):
        self.data = data

    @property
    def array(self):
        return self.data

    def reshape(self, *shape):
        return Variable(self.data.reshape(*shape))


def planes_to_bitboards(x):
    x = np.asarray(x.data if isinstance(x, Variable) else x)
    x = x.reshape(-1, 2, 64)
    opp = ((x[:, 0] != 0).astype(np.uint64) * _W).sum(axis=1, dtype=np.uint64)
    own = ((x[:, 1] != 0).astype(np.uint64) * _W).sum(axis=1, dtype=np.uint64)
    return own, opp


class _TrunkNet:
    kind = None
    default_file = None

    def __init__(self, device=0, precision=None):
        self.device, self.precision = device, precision
        self._tag = f"{type(self).__name__}#{next(_ids)}"
        self.slot = default_engine(device).alloc_slot(self._tag)   # owned until close() / garbage collection; raises when none is free
        self.params = None

    def close(self):
        if getattr(self, "slot", None) is not None:
            default_engine(self.device).free_slot(self.slot, self._tag)
            self.slot = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load(self, path):
        self.params = npz.read_npz(path)
        default_engine(self.device).load_net(self.slot, self.params, self.kind, owner=self._tag)
        return self

    def load_params(self, params):
        self.params = {k: np.asarray(v, np.float32) for k, v in params.items()}
        default_engine(self.device).load_net(self.slot, self.params, self.kind, owner=self._tag)
        return self


class SLPolicy(_TrunkNet):
    kind = npz.KIND_POLICY

    def __call__(self, x, probs=True):
        own, opp = planes_to_bitboards(x)
        out = default_engine(self.device).policy_forward_host(self.slot, own, opp, 1, probs=probs, precision=self.precision)
        return Variable(out)

    def forward_bitboards(self, p1, p2, color, probs=True):
        return default_engine(self.device).policy_forward_host(self.slot, p1, p2, color, probs=probs, precision=self.precision)


class Value(_TrunkNet):
    kind = npz.KIND_VALUE

    def __call__(self, x):
        own, opp = planes_to_bitboards(x)
        return Variable(default_engine(self.device).value_forward_host(self.slot, own, opp, 1, precision=self.precision))

    def forward_bitboards(self, p1, p2, color):
        return default_engine(self.device).value_forward_host(self.slot, p1, p2, color, precision=self.precision)


class RolloutPolicy:
    def __init__(self, device=0):
        self.device = device

    def load(self, path):
        default_engine(self.device).load_rollout_npz(path)
        return self

    def __call__(self, x):
        own, opp = planes_to_bitboards(x)
        logits = default_engine(self.device).rollout_logits_host(own, opp, 1).astype(np.float32)
        y = logits - logits.max(axis=1, keepdims=True)   # softmax is output formatting of the kernel's logits
        np.exp(y, out=y)
        y /= y.sum(axis=1, keepdims=True)
        return Variable(y)
