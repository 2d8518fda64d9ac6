"""torch (CPU, float64 by default) restatement of the REINFORCE update of src/train_rl.py:55-66 — TEST INFRASTRUCTURE ONLY.

    x    = stack([states == 1, states == 2])            (channel 0 = opponent, 1 = learner; recorded states are swapped)
    pred = SLPolicy(x)                                   softmax probabilities          network.py:34-47
    c    = softmax_cross_entropy(pred, y, reduce='no')   log-softmax applied to pred AGAIN (reference quirk)
    loss = mean(c * r);  loss.backward();  Adam + WeightDecay(5e-4)                     src/train_rl.py:24-26,61-66

Chainer's backward cannot be run here (Chainer is not installable and the numpy stand-in has no autograd), so this
restatement is the pin for K6: parity at this boundary is against torch autograd on the same graph — "parity unpinned" with
respect to Chainer itself.  Adam follows chainer.optimizers.Adam (AdamRule.update_core_cpu) and
chainer.optimizer_hooks.WeightDecay: g += rate * w; m += (1-b1)(g-m); v += (1-b2)(g*g-v);
w -= alpha * sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps).
"""
import numpy as np
import torch
import torch.nn.functional as F

KEYS = [k for i in range(1, 9) for k in (f"block{i}/conv/W", f"block{i}/conv/b")] + ["conv9/W", "bias10/b"]


def to_torch(params, dtype=torch.float64):
    return {k: torch.tensor(np.asarray(params[k]), dtype=dtype, requires_grad=True) for k in KEYS}


def forward(p, x, masks=None, keep=None):
    """masks (optional): 8 boolean arrays (M,C,8,8); when given, block i multiplies by masks[i-1] instead of applying ReLU — the
    derivative of ReLU at a pre-activation within rounding noise of 0 is ambiguous, and a checker for a lower-precision forward
    must take that forward's own on/off decisions as given (DESIGN.md "K6")."""
    h = x
    for i in range(1, 9):
        h = F.conv2d(h, p[f"block{i}/conv/W"], p[f"block{i}/conv/b"], padding=1)
        if keep is not None:
            keep.append(h.detach().numpy())
        h = F.relu(h) if masks is None else h * torch.tensor(np.asarray(masks[i - 1]), dtype=h.dtype)
    h = F.conv2d(h, p["conv9/W"]).reshape(-1, 64) + p["bias10/b"]
    return F.softmax(h, dim=1)


def loss_and_grad(params, states, actions, rewards, dtype=torch.float64, masks=None, keep=None):
    """states (M,8,8) in {0,1,2} as recorded by rl_self_play.Game (swapped). Returns (sum c*r, grads dict of d(sum c*r), pred).
    keep (optional list) receives the 8 pre-activation arrays."""
    p = to_torch(params, dtype)
    s = torch.tensor(np.asarray(states).reshape(-1, 8, 8))
    x = torch.stack([s == 1, s == 2], dim=1).to(dtype)
    pred = forward(p, x, masks, keep)
    c = F.cross_entropy(pred, torch.tensor(np.asarray(actions), dtype=torch.long), reduction="none")
    total = (c * torch.tensor(np.asarray(rewards), dtype=dtype)).sum()
    total.backward()
    return float(total.detach()), {k: v.grad.numpy() for k, v in p.items()}, pred.detach().numpy()


def flat(d):
    return np.concatenate([np.asarray(d[k], np.float64).reshape(-1) for k in KEYS])


def adam_step(w, g_mean, m, v, t, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4):
    """One optimizer.update() on flat float64 arrays; g_mean = gradient of the MEAN loss. Returns (w, m, v, t)."""
    t = t + 1
    g = g_mean + weight_decay * w
    m = m + (1 - beta1) * (g - m)
    v = v + (1 - beta2) * (g * g - v)
    lr = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return w - lr * m / (np.sqrt(v) + eps), m, v, t
