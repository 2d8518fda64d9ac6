"""The N > 1 path on CPU: two gloo ranks exercise the sharding helpers and the two collectives of the path
(iago_b200/parallel.py) — what bench.py and ReinforceTrainer run over NCCL on the GPUs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def fake_gradient(game_ids, n_params):
    """A deterministic stand-in for 'gradient of the games with these ids' (the real one needs a GPU)."""
    g = np.zeros(n_params + 2, np.float32)
    for gid in game_ids:
        rng = np.random.default_rng(int(gid))
        m = int(rng.integers(20, 34))
        g[:n_params] += rng.integers(-8, 9, n_params).astype(np.float32) * 0.125    # exactly representable: sums are order-free
        g[n_params] += float(rng.integers(-16, 17)) * 0.25
        g[n_params + 1] += m
    return g


def _worker(rank, world_size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from iago_b200 import parallel
    assert parallel.world() == (rank, world_size)
    n_params, n, steps = 1000, 8, 3
    res = []
    for step in range(steps):
        g0 = parallel.game_id0(step, rank, world_size, n)
        flat = torch.from_numpy(fake_gradient(range(g0, g0 + n), n_params))
        loss, count = parallel.mean_gradient_(flat, n_params)
        res.append((flat.numpy().copy(), loss, count))
    lo, hi = parallel.shard_range(1001, rank, world_size)
    counters = parallel.reduce_counters(dict(wins=rank + 1, games=hi - lo, plies=10 * (rank + 1)))
    t = parallel.all_reduce_max_(torch.tensor([float(rank)]))
    # root-parallel search of one game: each rank's visit counts, summed; lowest action wins ties; a lone pass child gives -1
    visits = np.zeros((3, 65), np.int32)
    visits[0, [19, 26, 37]] = [5 + rank, 7 - rank, 3]        # sums: 11, 13, 6 -> 26
    visits[1, [10, 20]] = [4 + rank, 5 - rank]                 # sums: 9, 9 -> tie -> 10
    visits[2, 64] = 8                                          # only the pass child
    vsum, best = parallel.root_parallel_moves(visits)
    # host-side control decisions of the REINFORCE loop (train_rl.train): rank 0's pick reaches every rank; the barrier helper
    choice = parallel.broadcast_object(f"model{rank + 7}.npz")
    parallel.barrier()
    # the supervised trainers' minibatch sharding: the epoch's permutation must be the same on every rank although the processes'
    # global np.random states differ (train_policy.train / train_value.train, world > 1 branch)
    np.random.seed(1234 + rank)
    seed, epoch, n_rec = 3, 2, 1001
    perm = np.random.RandomState((seed * 1000003 + epoch) & 0x7FFFFFFF).permutation(n_rec)
    lo_r, hi_r = parallel.shard_range(n_rec, rank, world_size)
    out[rank] = dict(res=res, counters=counters, tmax=float(t[0]), shard=(lo, hi), vsum=vsum.numpy(), best=best.tolist(), choice=choice,
                     perm_head=perm[:16].tolist(), mine=perm[lo_r:hi_r].tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_collectives():
    from iago_b200 import parallel
    world_size, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world_size, port, out), nprocs=world_size, join=True)
    n_params, n, steps = 1000, 8, 3
    for step in range(steps):
        # one process playing the union of the ranks' ids gets the same summed vector
        ids = range(step * world_size * n, (step + 1) * world_size * n)
        want = fake_gradient(ids, n_params)
        for r in range(world_size):
            flat, loss, count = out[r]["res"][step]
            assert (flat == want).all()
            assert count == want[n_params + 1] and abs(loss - want[n_params] / want[n_params + 1]) < 1e-12
    assert out[0]["counters"] == out[1]["counters"] == dict(wins=3, games=1001, plies=30)
    assert out[0]["tmax"] == out[1]["tmax"] == 1.0
    assert out[0]["shard"] == (0, 501) and out[1]["shard"] == (501, 1001)
    assert out[0]["choice"] == out[1]["choice"] == "model7.npz"
    assert out[0]["perm_head"] == out[1]["perm_head"]
    assert sorted(out[0]["mine"] + out[1]["mine"]) == list(range(1001))      # the two shards of one epoch cover every record exactly once
    for r in range(world_size):
        assert out[r]["best"] == [26, 10, -1]
        assert out[r]["vsum"][0, [19, 26, 37]].tolist() == [11, 13, 6] and int(out[r]["vsum"][2, 64]) == 16


def test_id_sharding_is_a_partition():
    from iago_b200 import parallel
    for world_size in (1, 2, 4, 8):
        seen = []
        for step in range(3):
            for r in range(world_size):
                g0 = parallel.game_id0(step, r, world_size, 16)
                seen.extend(range(g0, g0 + 16))
        assert sorted(seen) == list(range(3 * world_size * 16))
        cover = []
        for r in range(world_size):
            lo, hi = parallel.shard_range(1_000_000, r, world_size)
            cover.append((lo, hi))
        assert cover[0][0] == 0 and cover[-1][1] == 1_000_000 and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    assert parallel.world() == (0, 1)
    t = torch.ones(3)
    assert parallel.all_reduce_sum_(t) is t and (t == 1).all()     # no process group: no-op
