"""Chainer save_npz / load_npz key layout of the reference's weight files (network.py link names).

  SLPolicy : block{1..8}/conv/{W,b}, conv9/W, bias10/b                      (960,768 parameters)
  Value    : block{1..8}/conv/{W,b}, block9/conv/{W,b}, fc10/W, fc11/W      (970,049 parameters)
  Rollout  : conv1/W, bias2/b
rl_model.npz and models/RL_old/* carry a 'predictor/' prefix (saved through L.Classifier); it is stripped.
Only formatting lives here: the flat fp32 order is what iago_load_net documents (include/iago_b200.h).
"""
import numpy as np

KIND_POLICY, KIND_VALUE = 0, 1

TRUNK_KEYS = [k for i in range(1, 9) for k in (f"block{i}/conv/W", f"block{i}/conv/b")]
HEAD_KEYS = {KIND_POLICY: ["conv9/W", "bias10/b"],
             KIND_VALUE: ["block9/conv/W", "block9/conv/b", "fc10/W", "fc11/W"]}
N_PARAMS = {KIND_POLICY: 960768, KIND_VALUE: 970049}


def read_npz(path):
    with np.load(path) as z:
        out = {}
        for k in z.files:
            kk = k[len("predictor/"):] if k.startswith("predictor/") else k
            out[kk] = np.ascontiguousarray(z[k], dtype=np.float32)
    return out


def detect_kind(params):
    return KIND_VALUE if "fc10/W" in params else KIND_POLICY


def flatten(params, kind):
    keys = TRUNK_KEYS + HEAD_KEYS[kind]
    missing = [k for k in keys if k not in params]
    if missing:
        raise KeyError(f"weight archive lacks {missing}")
    flat = np.concatenate([np.asarray(params[k], np.float32).reshape(-1) for k in keys])
    assert flat.size == N_PARAMS[kind], flat.size
    return np.ascontiguousarray(flat)


def unflatten(flat, kind):
    shapes = {}
    cin = [2, 64, 128, 128, 128, 128, 128, 128]
    cout = [64, 128, 128, 128, 128, 128, 128, 128]
    for i in range(8):
        shapes[f"block{i + 1}/conv/W"] = (cout[i], cin[i], 3, 3)
        shapes[f"block{i + 1}/conv/b"] = (cout[i],)
    if kind == KIND_POLICY:
        shapes["conv9/W"] = (1, 128, 1, 1); shapes["bias10/b"] = (64,)
    else:
        shapes["block9/conv/W"] = (1, 128, 3, 3); shapes["block9/conv/b"] = (1,)
        shapes["fc10/W"] = (128, 64); shapes["fc11/W"] = (1, 128)
    out, o = {}, 0
    for k in TRUNK_KEYS + HEAD_KEYS[kind]:
        n = int(np.prod(shapes[k]))
        out[k] = np.asarray(flat[o:o + n], np.float32).reshape(shapes[k]).copy()
        o += n
    return out


def save_npz(path, params, prefix=""):
    """Chainer-compatible archive (np.savez_compressed with '/'-separated keys), loadable by the reference."""
    np.savez_compressed(path, **{prefix + k: v for k, v in params.items()})
