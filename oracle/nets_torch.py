"""torch restatement of the reference networks for LARGE batches — TEST INFRASTRUCTURE ONLY.

The same graph as oracle/nets.py (which follows /root/reference/network.py:5-96 and is pinned to the reference's own
outputs in tests/golden/nets.npz), evaluated with torch's conv2d on the CPU in float32 or float64: the numpy version
needs 4.7 ms per position, too slow for the >= 50,000-position comparison SURVEY.md 8d asks for.  It is itself pinned
to oracle/nets.py on the golden positions (tests/test_oracle_golden.py::test_torch_nets_equal_numpy_nets), so the chain
is reference outputs -> numpy restatement -> this file.  Chainer semantics restated: cross-correlation NCHW / OIHW with
zero padding (= torch.nn.functional.conv2d), Linear = x.reshape(N, -1) @ W.T.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(p, dtype):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in p.items()}


def _trunk(p, x):
    h = x
    for i in range(1, 9):
        h = F.relu(F.conv2d(h, p[f"block{i}/conv/W"], p[f"block{i}/conv/b"], padding=1))
    return h


@torch.no_grad()
def sl_logits(params, x, dtype=torch.float32, chunk=4096):
    """Pre-softmax logits (N,64) — network.py:34-46. params: dict of numpy arrays (oracle.nets.load_params)."""
    p = _t(params, dtype)
    out = []
    for i in range(0, len(x), chunk):
        h = _trunk(p, torch.from_numpy(np.ascontiguousarray(x[i:i + chunk])).to(dtype))
        h = F.conv2d(h, p["conv9/W"]).reshape(-1, 64) + p["bias10/b"].reshape(1, 64)
        out.append(h.to(torch.float64).numpy() if dtype == torch.float64 else h.numpy())
    return np.concatenate(out)


@torch.no_grad()
def value(params, x, dtype=torch.float32, chunk=4096):
    """network.py:83-96 at inference (dropout off, MCTS.py:86)."""
    p = _t(params, dtype)
    out = []
    for i in range(0, len(x), chunk):
        h = _trunk(p, torch.from_numpy(np.ascontiguousarray(x[i:i + chunk])).to(dtype))
        h = F.relu(F.conv2d(h, p["block9/conv/W"], p["block9/conv/b"], padding=1)).reshape(-1, 64)
        h = (h @ p["fc10/W"].T) @ p["fc11/W"].T
        out.append(h.reshape(-1).numpy())
    return np.concatenate(out)
