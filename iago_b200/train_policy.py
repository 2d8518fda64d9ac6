"""Supervised policy training — drop-in for /root/reference/train_policy.py:16-84 (same file under src/) on the GPU trainers.

    python -m iago_b200.train_policy --epoch 30 --policy sl        # or --policy rollout

Per epoch (train_policy.py:47-83): shuffle the training set (np.random.choice without replacement), minibatches of 4,096,
loss = F.softmax_cross_entropy(model(x), y) where model(x) is already softmax probabilities (the reference's double softmax,
kept), Adam (Chainer defaults) + WeightDecay(5e-4), then test loss / accuracy, a log line and Chainer-layout model / optimizer
archives.  States are "plays by 2" (load.py): channel 0 = (x == 1), channel 1 = (x == 2) = the mover (train_policy.py:10-11).
The SL policy step is the K6 gradient with reward 1 (ReinforceTrainer); the rollout policy has its own 82-parameter trainer.
With a process group the minibatch is sharded over the ranks and [gradient | loss | count] is all-reduced once per step.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import boards, npz, parallel
from ._lib import check
from .engine import default_engine
from .train_rl import N_PARAMS, ReinforceTrainer

MINIBATCH = 4096   # train_policy.py:44


def states_to_device(x, device):
    """(M,8,8) records in {0,1,2} -> (own = stones of 2, opp = stones of 1) int64 CUDA tensors."""
    p1, p2 = boards.to_bitboards(x)
    t = lambda a: torch.from_numpy(a.view(np.int64).copy()).to(device)
    return t(p2), t(p1)


class RolloutTrainer:
    """network.RolloutPolicy + optimizers.Adam + WeightDecay(5e-4) (train_policy.py:26-34)."""

    def __init__(self, conv1_W=None, bias2_b=None, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4, device=0, group=None, seed=0):
        self.eng = default_engine(device)
        self.lib, self.group = self.eng.lib, group
        self.hp = dict(alpha=alpha, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay)
        if conv1_W is None:   # Chainer's default initialisers: LeCunNormal for conv W (fan_in 18), zeros for the bias
            conv1_W = np.random.RandomState(seed).normal(0, np.sqrt(1.0 / 18), size=(1, 2, 3, 3))
            bias2_b = np.zeros(64)
        W = np.ascontiguousarray(conv1_W, np.float32).reshape(18)
        b = np.ascontiguousarray(bias2_b, np.float32).reshape(64)
        h = C.c_void_p()
        check(self.lib.iago_rollout_trainer_create(self.eng.ctx, W.ctypes.data, b.ctypes.data, C.byref(h)))
        self.h = h
        self.grad = torch.zeros(84, dtype=torch.float32, device=torch.device("cuda", device))

    def close(self):
        if getattr(self, "h", None):
            self.lib.iago_rollout_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gradient(self, own, opp, action, accumulate=False):
        if own.numel() == 0 and not accumulate:
            self.grad.zero_()   # an empty shard contributes nothing
        check(self.lib.iago_rollout_trainer_grad(self.h, C.c_void_p(own.data_ptr()), C.c_void_p(opp.data_ptr()), C.c_void_p(action.data_ptr()),
                                                 own.numel(), C.c_void_p(self.grad.data_ptr()), 1 if accumulate else 0, self.eng._stream(None)))

    def update(self):
        loss, count = parallel.mean_gradient_(self.grad, 82, self.group)
        if count <= 0:
            return 0.0, 0
        hp = self.hp
        check(self.lib.iago_rollout_trainer_adam_step(self.h, C.c_void_p(self.grad.data_ptr()), float(count), hp["alpha"], hp["beta1"], hp["beta2"],
                                                      hp["eps"], hp["weight_decay"], self.eng._stream(None)))
        return loss, int(count)

    def state(self):
        st = np.empty((3, 82), np.float32)
        t = C.c_int64()
        check(self.lib.iago_rollout_trainer_get_state(self.h, st.ctypes.data, C.byref(t)))
        return st[0].copy(), st[1].copy(), st[2].copy(), int(t.value)

    def params(self):
        p = self.state()[0]
        return {"conv1/W": p[:18].reshape(1, 2, 3, 3).copy(), "bias2/b": p[18:].copy()}

    def sync_engine(self):
        """Make the engine's rollout kernels play with the current parameters."""
        p = self.params()
        self.eng.load_rollout(p["conv1/W"], p["bias2/b"])

    def evaluate(self, own, opp, action):
        self.sync_engine()
        color = torch.ones(own.numel(), dtype=torch.uint8, device=own.device)
        logits = self.eng.rollout_logits(own, opp, color)      # p1 = mover's stones with colour 1
        return policy_metrics(self.eng, logits, action, is_logits=True)

    def save_model(self, path):
        npz.save_npz(path, self.params())

    def save_optimizer(self, path):
        p, m, v, t = self.state()
        d = {"t": np.array(t, np.int32), "epoch": np.array(0, np.int32)}
        for k, sl, shape in (("conv1/W", slice(0, 18), (1, 2, 3, 3)), ("bias2/b", slice(18, 82), (64,))):
            d[f"{k}/t"], d[f"{k}/m"], d[f"{k}/v"] = np.array(t, np.int32), m[sl].reshape(shape), v[sl].reshape(shape)
        np.savez_compressed(path, **d)


def policy_metrics(eng, values, action, is_logits=False):
    """(mean softmax_cross_entropy(pred, y), accuracy) as train_policy.py:69-70 computes them on the test set."""
    out = torch.zeros(2, dtype=torch.float32, device=values.device)
    n = action.numel()
    check(eng.lib.iago_policy_eval(eng.ctx, C.c_void_p(values.data_ptr()), 1 if is_logits else 0, C.c_void_p(action.data_ptr()), n,
                                   C.c_void_p(out.data_ptr()), eng._stream(None)))
    loss, hits = out.tolist()
    return loss / n, hits / n


class SLTrainer(ReinforceTrainer):
    """network.SLPolicy + Adam + WeightDecay(5e-4); one step = one minibatch of (state, action) records."""

    def step(self, own, opp, action):
        self.gradient(own, opp, action, torch.ones(own.numel(), dtype=torch.float32, device=own.device))
        return self.update()

    def evaluate(self, own, opp, action):
        color = torch.ones(own.numel(), dtype=torch.uint8, device=own.device)
        probs = torch.cat([self.eng.policy_forward(self.slot, own[i:i + 65536], opp[i:i + 65536], color[i:i + 65536], probs=True, precision=self.precision)
                           for i in range(0, own.numel(), 65536)])
        return policy_metrics(self.eng, probs, action)


def lecun_params(kind, seed=0):
    """Chainer's default initialisation of a fresh network.SLPolicy() / Value(): LeCunNormal weights, zero biases."""
    rs = np.random.RandomState(seed)
    shapes = npz.unflatten(np.zeros(npz.N_PARAMS[kind], np.float32), kind)
    out = {}
    for k, z in shapes.items():
        if k.endswith("/b"):
            out[k] = np.zeros_like(z)
        else:
            fan_in = int(np.prod(z.shape[1:]))
            out[k] = rs.normal(0, np.sqrt(1.0 / fan_in), size=z.shape).astype(np.float32)
    return out


def train(X_train, y_train, X_test, y_test, policy="sl", epochs=30, minibatch=MINIBATCH, model_path=None, optimizer_path=None, log=None,
          init=None, device=0, group=None, seed=0, on_epoch=None):
    """The loop of train_policy.py:47-83. Returns (trainer, [(test loss, test accuracy) per epoch])."""
    rank, world = parallel.world(group)
    dev = torch.device("cuda", device)
    if policy == "rollout":
        tr = RolloutTrainer(*(init or (None, None)), device=device, group=group, seed=seed)
    else:
        tr = SLTrainer(init or lecun_params(npz.KIND_POLICY, seed), max_positions=minibatch, device=device, group=group)
    tx_own, tx_opp = states_to_device(X_test, dev)
    ty = torch.from_numpy(np.asarray(y_test).astype(np.int8)).to(dev)
    X_train, y_train = np.asarray(X_train), np.asarray(y_train)
    n = y_train.shape[0]
    history = []
    for epoch in range(epochs):
        # One process: the reference's own draw from the global np.random (train_policy.py:49-50 / train_value.py:38-39).  Several ranks: the
        # permutation must be THE SAME on every rank (each takes its shard of every minibatch), so it comes from a generator seeded by
        # (seed, epoch) — the global np.random state of different processes is not synchronised.
        rands = np.random.choice(n, n, replace=False) if world == 1 else np.random.RandomState((seed * 1000003 + epoch) & 0x7FFFFFFF).permutation(n)
        X_train, y_train = X_train[rands], y_train[rands]
        for idx in range(0, n, minibatch):
            hi = min(idx + minibatch, n)
            lo_r, hi_r = parallel.shard_range(hi - idx, rank, world)
            own, opp = states_to_device(X_train[idx + lo_r:idx + hi_r], dev)
            act = torch.from_numpy(y_train[idx + lo_r:idx + hi_r].astype(np.int8)).to(dev)
            if policy == "rollout":
                tr.gradient(own, opp, act)
                tr.update()
            else:
                tr.step(own, opp, act)
        loss_test, test_acc = tr.evaluate(tx_own, tx_opp, ty)
        history.append((loss_test, test_acc))
        if rank == 0:
            if log:
                with open(log, "a") as f:
                    f.write(str(loss_test) + ", " + str(test_acc) + "\n")
            if model_path:
                tr.save_model(model_path)
            if optimizer_path:
                tr.save_optimizer(optimizer_path)
        if on_epoch:
            on_epoch(epoch, loss_test, test_acc)
    return tr, history


def main():
    import argparse
    ap = argparse.ArgumentParser(description="IaGo:")
    ap.add_argument("--epoch", "-e", type=int, default=30, help="Number of sweeps over the dataset to train")
    ap.add_argument("--policy", "-p", type=str, default="sl", help="Policy to train: sl or rollout")
    ap.add_argument("--gpu", "-g", type=int, default=0, help="GPU ID")
    args = ap.parse_args()
    if args.policy not in ("sl", "rollout"):
        print('Argument "--policy" is invalid. SLPolicy has been set by default.')
        args.policy = "sl"
    d = "../policy_data/npy/"
    name = "rollout" if args.policy == "rollout" else "sl"
    train(np.load(d + "states.npy"), np.load(d + "actions.npy"), np.load(d + "states_test.npy"), np.load(d + "actions_test.npy"),
          policy=args.policy, epochs=args.epoch, model_path=f"../models/{name}_model.npz", optimizer_path=f"../models/{name}_optimizer.npz",
          log=f"../log/{name}.txt", device=args.gpu,
          on_epoch=lambda e, l, a: print("\nepoch :", e, "  loss :", l, " accuracy:", a))


if __name__ == "__main__":
    main()
