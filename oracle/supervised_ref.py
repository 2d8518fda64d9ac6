"""torch (CPU, float64) restatements of the supervised training steps — TEST INFRASTRUCTURE ONLY.

  train_policy.py:56-64   loss = F.softmax_cross_entropy(model(x), y), model = SLPolicy or RolloutPolicy, both of which already
                          end in softmax (network.py:47,63): log-softmax is applied to the probabilities AGAIN (kept)
  train_value.py:50-56    loss = mean_squared_error(Value(x), y) with F.dropout(fc10(h), 0.4) active (network.py:94)
x = stack([state == 1, state == 2]) (train_policy.py:10-11, train_value.py:48).  The SL policy case is oracle/reinforce_ref.py with
reward 1.  Chainer's backward cannot be run here (not installable): parity is against torch autograd on the same graph
("parity unpinned" w.r.t. Chainer).  Sums, not means, are returned; the trainers divide by the record count.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .reinforce_ref import KEYS as POLICY_KEYS  # noqa: F401  (re-exported for the tests)

VALUE_KEYS = [k for i in range(1, 9) for k in (f"block{i}/conv/W", f"block{i}/conv/b")] + ["block9/conv/W", "block9/conv/b", "fc10/W", "fc11/W"]


def planes(states, dtype=torch.float64):
    s = torch.tensor(np.asarray(states).reshape(-1, 8, 8))
    return torch.stack([s == 1, s == 2], dim=1).to(dtype)


def rollout_loss_and_grad(W, b, states, actions):
    """Returns (sum of CE, dW (1,2,3,3), db (64,), probabilities)."""
    W = torch.tensor(np.asarray(W, np.float64).reshape(1, 2, 3, 3), requires_grad=True)
    b = torch.tensor(np.asarray(b, np.float64).reshape(64), requires_grad=True)
    pred = F.softmax(F.conv2d(planes(states), W, padding=1).reshape(-1, 64) + b, dim=1)
    total = F.cross_entropy(pred, torch.tensor(np.asarray(actions), dtype=torch.long), reduction="sum")
    total.backward()
    return float(total.detach()), W.grad.numpy(), b.grad.numpy(), pred.detach().numpy()


def value_loss_and_grad(params, states, targets, drop_mask=None, ratio=0.4, relu_masks=None, keep=None):
    """drop_mask (M,128) of {0,1} (1 = kept) or None for evaluation; relu_masks as in reinforce_ref.forward (8 arrays (M,C,8,8)).
    Returns (sum (v - y)^2, grads dict, v)."""
    p = {k: torch.tensor(np.asarray(params[k], np.float64), requires_grad=True) for k in VALUE_KEYS}
    h = planes(states)
    for i in range(1, 9):
        h = F.conv2d(h, p[f"block{i}/conv/W"], p[f"block{i}/conv/b"], padding=1)
        if keep is not None:
            keep.append(h.detach().numpy())
        h = F.relu(h) if relu_masks is None else h * torch.tensor(np.asarray(relu_masks[i - 1]), dtype=h.dtype)
    h = F.relu(F.conv2d(h, p["block9/conv/W"], p["block9/conv/b"], padding=1)).reshape(-1, 64)
    u = h @ p["fc10/W"].T
    if drop_mask is not None:
        u = u * torch.tensor(np.asarray(drop_mask), dtype=u.dtype) / (1.0 - ratio)
    v = (u @ p["fc11/W"].T).reshape(-1)
    total = ((v - torch.tensor(np.asarray(targets, np.float64))) ** 2).sum()
    total.backward()
    return float(total.detach()), {k: t.grad.numpy() for k, t in p.items()}, v.detach().numpy()
