"""Value-network training — drop-in for /root/reference/train_value.py:9-70 on the GPU trainer.

    python -m iago_b200.train_value --epoch 20

Per epoch (train_value.py:36-69): shuffle, minibatches of 4,096, loss = mean_squared_error(model(x), y) with the training-mode
dropout of network.py:94 (ratio 0.4, between fc10 and fc11), Adam (Chainer defaults) + WeightDecay(5e-4), then the test loss
(dropout off), a log line and Chainer-layout archives.  x: channel 0 = (state == 1), channel 1 = (state == 2) (train_value.py:24,48):
the records of value_self_play / gen_value_data.py hold the mover as 2.
"""
import ctypes as C

import numpy as np
import torch

from . import npz, parallel
from ._lib import check
from .engine import default_engine
from .train_policy import MINIBATCH, lecun_params, states_to_device

N_PARAMS = npz.N_PARAMS[npz.KIND_VALUE]


class ValueTrainer:
    def __init__(self, params=None, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4, max_positions=MINIBATCH, device=0,
                 slot=None, precision=3, group=None, tensor_cores=True, dropout=0.4, seed=0):
        self.eng = default_engine(device)
        self.lib = self.eng.lib
        self._tag = f"ValueTrainer@{id(self):x}"
        self._own_slot = slot is None   # None: a free slot from the engine's allocator (released by close()); a number: the caller's slot
        if slot is None:
            slot = self.eng.alloc_slot(self._tag)
        self.device, self.slot, self.precision, self.group = device, slot, precision, group
        self.hp = dict(alpha=alpha, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay)
        self.dropout, self.seed, self.records = float(dropout), int(seed), 0
        if params is None:
            params = lecun_params(npz.KIND_VALUE, seed)
        elif isinstance(params, (str, bytes)) or hasattr(params, "__fspath__"):
            params = npz.read_npz(params)
        flat = npz.flatten(params, npz.KIND_VALUE)
        h = C.c_void_p()
        check(self.lib.iago_trainer_create(self.eng.ctx, npz.KIND_VALUE, flat.ctypes.data, flat.size, int(max_positions), C.byref(h)))
        self.h, self.max_positions = h, int(max_positions)
        check(self.lib.iago_reinforce_set_option(self.h, 1 if tensor_cores else 0))
        self.grad = torch.zeros(N_PARAMS + 2, dtype=torch.float32, device=torch.device("cuda", device))
        self.sync_slot()

    def close(self):
        if getattr(self, "h", None):
            self.lib.iago_reinforce_destroy(self.h)
            self.h = None
        if getattr(self, "_own_slot", False) and getattr(self, "slot", None) is not None:
            self.eng.free_slot(self.slot, self._tag)
            self.slot = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync_slot(self):
        check(self.lib.iago_reinforce_sync_slot(self.h, int(self.slot)))

    def gradient(self, own, opp, target, accumulate=False, dropout=None, want_pred=False, want_mask=False, position_id0=None):
        m = own.numel()
        if m == 0 and not accumulate:
            self.grad.zero_()   # an empty shard contributes nothing (not the previous step's all-reduced vector)
        ratio = self.dropout if dropout is None else float(dropout)
        pred = torch.empty(m, dtype=torch.float32, device=own.device) if want_pred else None
        mask = torch.empty((m, 128), dtype=torch.uint8, device=own.device) if want_mask else None
        pid0 = self.records if position_id0 is None else int(position_id0)
        for lo in range(0, m, self.max_positions):
            hi = min(m, lo + self.max_positions)
            sl = slice(lo, hi)
            check(self.lib.iago_value_grad(self.h, C.c_void_p(own[sl].data_ptr()), C.c_void_p(opp[sl].data_ptr()), C.c_void_p(target[sl].data_ptr()),
                                           hi - lo, C.c_void_p(self.grad.data_ptr()), 1 if (accumulate or lo > 0) else 0, ratio, self.seed, pid0 + lo,
                                           C.c_void_p(pred[sl].data_ptr()) if want_pred else None,
                                           C.c_void_p(mask[sl].data_ptr()) if want_mask else None, self.eng._stream(None)))
        self.records += m
        return pred, mask

    def update(self):
        loss, count = parallel.mean_gradient_(self.grad, N_PARAMS, self.group)
        if count <= 0:
            return 0.0, 0
        hp = self.hp
        check(self.lib.iago_reinforce_adam_step(self.h, C.c_void_p(self.grad.data_ptr()), float(count), hp["alpha"], hp["beta1"], hp["beta2"],
                                                hp["eps"], hp["weight_decay"], self.eng._stream(None)))
        self.sync_slot()
        return loss, int(count)

    def step(self, own, opp, target):
        self.gradient(own, opp, target)
        return self.update()

    def evaluate(self, own, opp, target):
        """mean_squared_error(model(x), y) with dropout off (train_value.py:58-61)."""
        color = torch.ones(own.numel(), dtype=torch.uint8, device=own.device)
        pred = torch.cat([self.eng.value_forward(self.slot, own[i:i + 65536], opp[i:i + 65536], color[i:i + 65536], precision=self.precision)
                          for i in range(0, own.numel(), 65536)])
        out = torch.zeros(2, dtype=torch.float32, device=own.device)
        check(self.lib.iago_value_eval(self.eng.ctx, C.c_void_p(pred.data_ptr()), C.c_void_p(target.data_ptr()), own.numel(),
                                       C.c_void_p(out.data_ptr()), self.eng._stream(None)))
        return float(out[0]) / own.numel()

    def state(self):
        p, m, v = (np.empty(N_PARAMS, np.float32) for _ in range(3))
        t = C.c_int64()
        check(self.lib.iago_reinforce_get_state(self.h, p.ctypes.data, m.ctypes.data, v.ctypes.data, C.byref(t)))
        return p, m, v, int(t.value)

    def params(self):
        return npz.unflatten(self.state()[0], npz.KIND_VALUE)

    def save_model(self, path):
        npz.save_npz(path, self.params())

    def save_optimizer(self, path):
        p, m, v, t = self.state()
        M, V = npz.unflatten(m, npz.KIND_VALUE), npz.unflatten(v, npz.KIND_VALUE)
        d = {"t": np.array(t, np.int32), "epoch": np.array(0, np.int32)}
        for k in M:
            d[f"{k}/t"], d[f"{k}/m"], d[f"{k}/v"] = np.array(t, np.int32), M[k], V[k]
        np.savez_compressed(path, **d)


def train(train_x, train_y, test_x, test_y, epochs=20, minibatch=MINIBATCH, model_path=None, optimizer_path=None, log=None, init=None,
          device=0, group=None, seed=0, on_epoch=None):
    """The loop of train_value.py:36-69. Returns (trainer, [test loss per epoch])."""
    rank, world = parallel.world(group)
    dev = torch.device("cuda", device)
    tr = ValueTrainer(init, max_positions=minibatch, device=device, group=group, seed=seed)
    tx_own, tx_opp = states_to_device(test_x, dev)
    ty = torch.from_numpy(np.asarray(test_y).astype(np.float32)).to(dev)
    train_x, train_y = np.asarray(train_x), np.asarray(train_y)
    n = train_y.shape[0]
    history = []
    for epoch in range(epochs):
        # One process: the reference's own draw from the global np.random (train_policy.py:49-50 / train_value.py:38-39).  Several ranks: the
        # permutation must be THE SAME on every rank (each takes its shard of every minibatch), so it comes from a generator seeded by
        # (seed, epoch) — the global np.random state of different processes is not synchronised.
        rands = np.random.choice(n, n, replace=False) if world == 1 else np.random.RandomState((seed * 1000003 + epoch) & 0x7FFFFFFF).permutation(n)
        train_x, train_y = train_x[rands], train_y[rands]
        for idx in range(0, n, minibatch):
            hi = min(idx + minibatch, n)
            lo_r, hi_r = parallel.shard_range(hi - idx, rank, world)
            own, opp = states_to_device(train_x[idx + lo_r:idx + hi_r], dev)
            y = torch.from_numpy(train_y[idx + lo_r:idx + hi_r].astype(np.float32)).to(dev)
            tr.gradient(own, opp, y, position_id0=epoch * n + idx + lo_r)
            tr.update()
        test_loss = tr.evaluate(tx_own, tx_opp, ty)
        history.append(test_loss)
        if rank == 0:
            if log:
                with open(log, "a") as f:
                    f.write(("%.6f" % test_loss)[:6] + ", \n")     # train_value.py:65 slices the printed Variable to 6 characters
            if model_path:
                tr.save_model(model_path)
            if optimizer_path:
                tr.save_optimizer(optimizer_path)
        if on_epoch:
            on_epoch(epoch, test_loss)
    return tr, history


def main():
    import argparse
    ap = argparse.ArgumentParser(description="IaGo:")
    ap.add_argument("--epoch", "-e", type=int, default=20, help="Number of sweeps over the dataset to train")
    ap.add_argument("--gpu", "-g", type=int, default=0, help="GPU ID to be used")
    args = ap.parse_args()
    d = "./value_data/npy/"
    train(np.load(d + "states.npy"), np.load(d + "results.npy"), np.load(d + "states_test.npy"), np.load(d + "results_test.npy"),
          epochs=args.epoch, model_path="./models/value_model.npz", optimizer_path="./models/value_optimizer.npz", log="./log_value.txt",
          device=args.gpu, on_epoch=lambda e, l: print("\nepoch :", e, "  loss :", l))


if __name__ == "__main__":
    main()
