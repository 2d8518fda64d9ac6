"""REINFORCE self-play training — the update of /root/reference/src/train_rl.py:39-66 on the GPU, sharded over ranks.

    trainer = ReinforceTrainer("models/RL/model2.npz", alpha=1e-3)          # src/train_rl.py:22-26
    stats = trainer.train_set(opponent, n_games=64)                          # one `while` body of src/train_rl.py:32-66

One set = 2N games of rl_self_play.Game(model1, model2) — odd-numbered games start from a 'head/tail switched' opening
(an extra, un-flipped colour-2 stone on one of (2,4),(3,5),(4,2),(5,3), src/train_rl.py:43-46) — then ONE update:
loss = mean(softmax_cross_entropy(model1(x), y, reduce='no') * r), Adam + WeightDecay(5e-4).

Data parallel: every rank plays its own games (global game ids key the RNG) and computes the gradient of SUM c*r over its
own positions; the flat vector [gradient | SUM c*r | position count] (3.84 MB) is all-reduced (NCCL, sum) and every rank
applies the identical Adam step to its replica — exactly the update one process would make on the union of the games.
"""
import ctypes as C

import numpy as np
import torch

from . import boards, npz, parallel
from ._lib import check
from .engine import Rng, STREAM_SELFPLAY, default_engine

N_PARAMS = npz.N_PARAMS[npz.KIND_POLICY]
SWITCH_CELLS = [2 * 8 + 4, 3 * 8 + 5, 4 * 8 + 2, 5 * 8 + 3]   # src/train_rl.py:45


class ReinforceTrainer:
    def __init__(self, params, alpha=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4, max_positions=8192,
                 device=0, slot=None, precision=3, group=None, tensor_cores=True, comm=None, rec_cap=40):
        self.eng = default_engine(device)
        self.lib = self.eng.lib
        self._tag = f"ReinforceTrainer@{id(self):x}"
        self._own_slot = slot is None   # None: a free slot from the engine's allocator (released by close()); a number: the caller's slot
        if slot is None:
            slot = self.eng.alloc_slot(self._tag)
        self.device, self.slot, self.precision, self.group = device, slot, precision, group
        self.hp = dict(alpha=alpha, beta1=beta1, beta2=beta2, eps=eps, weight_decay=weight_decay)
        if isinstance(params, (str, bytes)) or hasattr(params, "__fspath__"):
            params = npz.read_npz(params)
        flat = npz.flatten(params, npz.KIND_POLICY)
        h = C.c_void_p()
        check(self.lib.iago_reinforce_create(self.eng.ctx, flat.ctypes.data, flat.size, int(max_positions), C.byref(h)))
        self.h, self.max_positions = h, int(max_positions)
        check(self.lib.iago_reinforce_set_option(self.h, 1 if tensor_cores else 0))
        self.grad = torch.zeros(N_PARAMS + 2, dtype=torch.float32, device=torch.device("cuda", device))
        # `comm` (parallel.Communicator): the gradient all-reduce runs over NCCL through the library, enqueued on the stream; without it a
        # torch.distributed group (gloo in the CPU tests) is the transport, and one process needs neither
        self.comm, self.rec_cap = comm, int(rec_cap)
        self._set_bufs = None        # compacted records of a set (device) + their pinned counters
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

    # ---- the learner as a playable model (rl_self_play.play_games wants .slot / .precision)
    def sync_slot(self):
        check(self.lib.iago_reinforce_sync_slot(self.h, int(self.slot)))

    # ---- gradient of SUM c*r over a batch of recorded decisions (device tensors)
    def gradient(self, own, opp, action, reward, accumulate=False, want_probs=False):
        m = own.numel()
        if m == 0 and not accumulate:
            self.grad.zero_()   # an empty shard (fewer records than ranks) contributes nothing: not the previous step's all-reduced vector
        probs = torch.empty((m, 64), dtype=torch.float32, device=own.device) if want_probs else None
        for lo in range(0, m, self.max_positions):
            hi = min(m, lo + self.max_positions)
            sl = slice(lo, hi)
            check(self.lib.iago_reinforce_grad(self.h, C.c_void_p(own[sl].data_ptr()), C.c_void_p(opp[sl].data_ptr()),
                                               C.c_void_p(action[sl].data_ptr()), C.c_void_p(reward[sl].data_ptr()), hi - lo,
                                               C.c_void_p(self.grad.data_ptr()), 1 if (accumulate or lo > 0) else 0,
                                               C.c_void_p(probs[sl].data_ptr()) if want_probs else None, self.eng._stream(None)))
        return probs

    def update(self, want_stats=True):
        """optimizer.update(): all-reduce [gradient | loss numerator | count] over the ranks (the only collective of the training
        path), WeightDecay hook, Adam, refresh the playing slot.  Everything is enqueued on the stream: the Adam kernel reads the
        reduced count on the device and the slot refresh follows it.  Returns (mean loss, positions) — one 8-byte read after the
        whole update has been enqueued — or None with want_stats=False (no host synchronisation at all)."""
        rank, world = parallel.world(self.group)
        if self.comm is not None:
            self.comm.all_reduce_sum_(self.grad)
        elif world > 1:
            parallel.all_reduce_sum_(self.grad, self.group)      # torch.distributed transport (gloo in the CPU tests)
        hp = self.hp
        stream = self.eng._stream(None)
        check(self.lib.iago_reinforce_adam_step_dev(self.h, C.c_void_p(self.grad.data_ptr()), hp["alpha"], hp["beta1"], hp["beta2"],
                                                    hp["eps"], hp["weight_decay"], stream))
        check(self.lib.iago_reinforce_sync_slot_async(self.h, int(self.slot), stream))
        if not want_stats:
            return None
        num, count = self.grad[N_PARAMS:N_PARAMS + 2].tolist()
        return (num / count if count > 0 else 0.0), int(count)

    # ---- one set of games + one update
    def play_set(self, opponent, n_games, seed=0, game_id0=0):
        """2N games vs `opponent` (an SLPolicy with a resident slot): openings, games and the flattening of the learner's records all
        run on the device (iago_reinforce_openings / iago_selfplay / iago_reinforce_compact); one 12-byte read tells the host how many
        records there are.  Returns the flattened learner decisions (device views) and the set's counters."""
        dev = self.grad.device
        stream = self.eng._stream(None)
        cap = self.rec_cap
        if self._set_bufs is None or self._set_bufs["n"] < n_games:
            self._set_bufs = dict(n=n_games, i1=torch.empty(n_games, dtype=torch.int64, device=dev), i2=torch.empty(n_games, dtype=torch.int64, device=dev),
                                  own=torch.empty(n_games * cap, dtype=torch.int64, device=dev), opp=torch.empty(n_games * cap, dtype=torch.int64, device=dev),
                                  action=torch.empty(n_games * cap, dtype=torch.int8, device=dev), reward=torch.empty(n_games * cap, dtype=torch.float32, device=dev),
                                  count=torch.zeros(3, dtype=torch.int32, device=dev), count_host=torch.zeros(3, dtype=torch.int32).pin_memory())
        b = self._set_bufs
        i1, i2 = b["i1"][:n_games], b["i2"][:n_games]
        check(self.lib.iago_reinforce_openings(self.eng.ctx, n_games, int(seed) & (2**64 - 1), int(game_id0), C.c_void_p(i1.data_ptr()),
                                               C.c_void_p(i2.data_ptr()), stream))
        out = self.eng.selfplay(self.slot, opponent.slot, n_games, i1, i2, greedy=False, precision=None, rec_cap=cap,
                                rng=Rng.philox(seed=seed, game_id0=game_id0, stream_id=STREAM_SELFPLAY))
        p = lambda t: C.c_void_p(t.data_ptr())
        check(self.lib.iago_reinforce_compact(self.eng.ctx, n_games, cap, p(out["rec_own"]), p(out["rec_opp"]), p(out["rec_action"]), p(out["n_rec"]),
                                              p(out["result"]), p(b["own"]), p(b["opp"]), p(b["action"]), p(b["reward"]), p(b["count"]), stream))
        b["count_host"].copy_(b["count"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        m, wins, longest = (int(v) for v in b["count_host"])
        if longest > cap:
            raise RuntimeError(f"a game recorded {longest} learner decisions, more than rec_cap = {cap}")
        return dict(own=b["own"][:m], opp=b["opp"][:m], action=b["action"][:m], reward=b["reward"][:m], wins=wins, games=n_games)

    def train_set(self, opponent, n_games=64, seed=0, game_id0=0, want_stats=True):
        d = self.play_set(opponent, n_games, seed=seed, game_id0=game_id0)
        self.gradient(d["own"], d["opp"], d["action"], d["reward"])
        st = self.update(want_stats)
        if st is None:
            return dict(rate=d["wins"] / d["games"])
        return dict(rate=d["wins"] / d["games"], loss=st[0], positions=st[1])

    # ---- checkpoints in the reference's formats (serializers.save_npz of the model and of the optimizer)
    def state(self):
        p, m, v = (np.empty(N_PARAMS, np.float32) for _ in range(3))
        t = C.c_int64()
        check(self.lib.iago_reinforce_get_state(self.h, p.ctypes.data, m.ctypes.data, v.ctypes.data, C.byref(t)))
        return p, m, v, int(t.value)

    def params(self):
        return npz.unflatten(self.state()[0], npz.KIND_POLICY)

    def load_state(self, params=None, adam_m=None, adam_v=None, step=-1):
        a = lambda x: None if x is None else np.ascontiguousarray(x, np.float32).ctypes.data
        keep = [None if x is None else np.ascontiguousarray(x, np.float32) for x in (params, adam_m, adam_v)]
        check(self.lib.iago_reinforce_set_state(self.h, *[None if k is None else k.ctypes.data for k in keep], int(step)))
        self.sync_slot()

    def save_model(self, path, prefix=""):
        npz.save_npz(path, self.params(), prefix=prefix)

    def save_optimizer(self, path):
        """Chainer optimizer archive layout (as models/rollout_optimizer.npz): 't', 'epoch' and per parameter '<path>/t', '/m', '/v'."""
        p, m, v, t = self.state()
        M, V = npz.unflatten(m, npz.KIND_POLICY), npz.unflatten(v, npz.KIND_POLICY)
        d = {"t": np.array(t, np.int32), "epoch": np.array(0, np.int32)}
        for k in M:
            d[f"{k}/t"], d[f"{k}/m"], d[f"{k}/v"] = np.array(t, np.int32), M[k], V[k]
        np.savez_compressed(path, **d)


class PoolSchedule:
    """The bookkeeping of src/train_rl.py:29-81: `models` counts the snapshots in the opponent pool, `cnt` the sets won
    (rate > 0.5) since the last snapshot; a snapshot is taken when cnt > 4*sqrt(models) and rate > 0.6 (:73-79); training
    stops when rate < 0.2 (:80-81) or models > 20 (:32)."""

    def __init__(self, models=1, max_models=20):
        self.models, self.cnt, self.max_models = int(models), 0, int(max_models)

    def running(self):
        return self.models <= self.max_models

    def step(self, rate):
        """Returns (snapshot_index or None, stop)."""
        snap = None
        if rate > 0.5:
            self.cnt += 1
        if self.cnt > 4 * np.sqrt(self.models) and rate > 0.6:
            snap = self.models
            self.models += 1
            self.cnt = 0
        return snap, rate < 0.2


def train(model_dir, start="model2.npz", models=1, n_games=64, alpha=1e-3, max_sets=None, seed=0, log=None, device=0,
          group=None, on_set=None):
    """The outer loop of src/train_rl.py:22-81 on the GPU trainer: random opponent from model_dir/*.npz per set
    (np.random.choice over the glob, :35-37), one update per set, win-rate log line, pool snapshots
    model<k>.npz + optimizers/<k>.npz in the reference's archive layouts, early stop. Returns the list of per-set stats."""
    import glob
    import os
    from . import network, parallel
    rank, world = parallel.world(group)
    trainer = ReinforceTrainer(os.path.join(model_dir, start), alpha=alpha, device=device, group=group)
    opt = os.path.join(model_dir, "optimizers", os.path.splitext(start)[0].replace("model", "") + ".npz")
    if os.path.isfile(opt):   # src/train_rl.py:27 resumes the optimizer when its archive exists
        with np.load(opt) as z:
            keys = npz.TRUNK_KEYS + npz.HEAD_KEYS[npz.KIND_POLICY]
            m = np.concatenate([np.asarray(z[k + "/m"], np.float32).reshape(-1) for k in keys])
            v = np.concatenate([np.asarray(z[k + "/v"], np.float32).reshape(-1) for k in keys])
            trainer.load_state(None, m, v, int(z["t"]))
    sched = PoolSchedule(models)
    rng = np.random.RandomState(seed)     # the same draw on every rank: all ranks face the same opponent
    opponent = network.SLPolicy(device=device)
    history, s = [], 0
    while sched.running() and (max_sets is None or s < max_sets):
        # Rank 0 lists the pool and picks; the choice is broadcast so that every rank faces the same opponent even while rank 0 is
        # adding snapshots to the directory (a snapshot appears atomically — os.replace below — and a barrier follows it)
        pool = sorted(f for f in glob.glob(os.path.join(model_dir, "*.npz")) if not f.endswith(".tmp.npz"))
        path = parallel.broadcast_object(pool[rng.randint(len(pool))], group)
        opponent.load(path)
        st = trainer.play_set(opponent, n_games, seed=seed, game_id0=parallel.game_id0(s, rank, world, n_games))
        trainer.gradient(st["own"], st["opp"], st["action"], st["reward"])
        loss, count = trainer.update()
        wins = parallel.reduce_counters(dict(wins=st["wins"], games=st["games"]), device=trainer.grad.device, group=group)
        rate = wins["wins"] / wins["games"]
        snap, stop = sched.step(rate)
        rec = dict(set=s, opponent=os.path.basename(path), rate=rate, loss=loss, positions=count, models=sched.models, snapshot=snap)
        history.append(rec)
        if rank == 0:
            if log:
                with open(log, "a") as f:
                    f.write(str(rate) + ", \n")          # src/train_rl.py:69-70
            if snap is not None:
                os.makedirs(os.path.join(model_dir, "optimizers"), exist_ok=True)
                for writer, dst in ((trainer.save_model, os.path.join(model_dir, f"model{snap}.npz")),
                                    (trainer.save_optimizer, os.path.join(model_dir, "optimizers", f"{snap}.npz"))):
                    tmp = dst[:-4] + ".tmp.npz"          # never a partially written archive under its final name
                    writer(tmp)
                    os.replace(tmp, dst)
        if snap is not None:
            parallel.barrier(group)                      # the new pool member exists before any rank lists the pool again
        if on_set:
            on_set(rec)
        s += 1
        if stop:
            break
    return history


def main():
    import argparse
    ap = argparse.ArgumentParser(description="IaGo: REINFORCE self-play on B200 (src/train_rl.py)")
    ap.add_argument("--models", "-m", type=int, default=1, help="Number of trained models")
    ap.add_argument("--set", "-s", type=int, default=1000, help="Number of game sets played to train")
    ap.add_argument("--dir", default="../models/RL")
    ap.add_argument("--games", type=int, default=64, help="games per set (2N, N = 32 in the reference)")
    args = ap.parse_args()
    train(args.dir, models=args.models, n_games=args.games, max_sets=args.set, log="../log/rl.txt",
          on_set=lambda r: print("Models:" + str(r["models"]) + ", Result:" + str(r["rate"]) + ", Loss:" + str(r["loss"])))


if __name__ == "__main__":
    main()
