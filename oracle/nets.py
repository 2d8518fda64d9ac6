"""numpy restatement of the reference networks — TEST INFRASTRUCTURE ONLY.

Follows /root/reference/network.py:
  Block        network.py:5-13    3x3 conv pad 1 + bias + ReLU
  SLPolicy     network.py:15-47   8 blocks (2->64->128 x7) -> conv9 1x1 128->1 (no bias) -> +bias10.b[64] -> softmax
  RolloutPolicy network.py:49-64  conv1 2->1 3x3 pad 1 (no bias) -> +bias2.b[64] -> softmax
  Value        network.py:66-96   same 8-block trunk -> block9 (3x3 128->1 +b +ReLU) -> fc10 (64->128, no bias)
                                  -> dropout (identity at inference, MCTS.py:86) -> fc11 (128->1, no bias)
and the input encoding game.py:167-174 (channel 0 = opponent stones, channel 1 = mover's stones).

Chainer semantics restated (third-party, un-vendored, un-pinned — SURVEY.md §8c; parity at this boundary
is UNPINNED apart from the golden outputs produced with the numpy stand-in in oracle/chainer_shim):
cross-correlation NCHW/OIHW with zero padding; Linear = x.reshape(N,-1) @ W.T; softmax = exp(x-max)/sum.

`dtype=np.float64` gives the error yard-stick; `dtype=np.float32` the reference-precision forward.
"""
import numpy as np


def load_params(path, dtype=np.float32):
    """Chainer save_npz layout; strips an optional 'predictor/' prefix (rl_model.npz, RL_old/*)."""
    with np.load(path) as z:
        out = {}
        for k in z.files:
            kk = k[len("predictor/"):] if k.startswith("predictor/") else k
            out[kk] = np.ascontiguousarray(z[k]).astype(dtype)
    return out


def planes_from_state(states, colors, dtype=np.float32):
    """game.py:167-174 batched: states (N,8,8) in {0,1,2}, colors (N,) -> (N,2,8,8)."""
    s = np.asarray(states).reshape(-1, 8, 8)
    c = np.broadcast_to(np.asarray(colors), (s.shape[0],)).reshape(-1, 1, 1)
    mover = (s == c)
    opp = (s == (3 - c))
    return np.stack([opp, mover], axis=1).astype(dtype)


def planes_from_bitboards(own, opp, dtype=np.float32):
    """own = mover's stones, opp = opponent's; bit k <-> cell k."""
    sh = np.arange(64, dtype=np.uint64)
    o = ((np.asarray(own, np.uint64).reshape(-1, 1) >> sh) & np.uint64(1)).reshape(-1, 8, 8)
    p = ((np.asarray(opp, np.uint64).reshape(-1, 1) >> sh) & np.uint64(1)).reshape(-1, 8, 8)
    return np.stack([p, o], axis=1).astype(dtype)


def conv2d(x, W, b=None, pad=1):
    n, c, h, w = x.shape
    o, _, k, _ = W.shape
    if k == 1:
        y = np.einsum("nchw,oc->nohw", x, W[:, :, 0, 0], optimize=True)
    else:
        xp = np.zeros((n, c, h + 2 * pad, w + 2 * pad), x.dtype)
        xp[:, :, pad:pad + h, pad:pad + w] = x
        col = np.empty((n, h, w, c, k, k), x.dtype)
        for i in range(k):
            for j in range(k):
                col[:, :, :, :, i, j] = xp[:, :, i:i + h, j:j + w].transpose(0, 2, 3, 1)
        y = col.reshape(n * h * w, c * k * k) @ W.reshape(o, c * k * k).T
        y = y.reshape(n, h, w, o).transpose(0, 3, 1, 2)
    if b is not None:
        y = y + b.reshape(1, -1, 1, 1)
    return np.ascontiguousarray(y.astype(x.dtype, copy=False))


def softmax(x):
    y = x - x.max(axis=1, keepdims=True)
    np.exp(y, out=y)
    y /= y.sum(axis=1, keepdims=True)
    return y


def trunk(p, x, upto=8, collect=None):
    h = x
    for i in range(1, upto + 1):
        h = np.maximum(conv2d(h, p[f"block{i}/conv/W"], p[f"block{i}/conv/b"]), 0)
        if collect is not None:
            collect.append(h)
    return h


def sl_logits(p, x):
    """Pre-softmax logits (N,64) — network.py:34-46."""
    h = trunk(p, x)
    h = conv2d(h, p["conv9/W"], None, pad=0).reshape(-1, 64)
    return h + p["bias10/b"].reshape(1, 64)


def sl_policy(p, x):
    return softmax(sl_logits(p, x))


def rollout_logits(p, x):
    h = conv2d(x, p["conv1/W"], None).reshape(-1, 64)
    return h + p["bias2/b"].reshape(1, 64)


def rollout_policy(p, x):
    return softmax(rollout_logits(p, x))


def value(p, x):
    h = trunk(p, x)
    h = np.maximum(conv2d(h, p["block9/conv/W"], p["block9/conv/b"]), 0).reshape(-1, 64)
    h = h @ p["fc10/W"].T
    h = h @ p["fc11/W"].T
    return h.reshape(-1)
