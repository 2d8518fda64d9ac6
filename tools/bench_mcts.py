#!/usr/bin/env python
"""PV-MCTS playouts/s (BASELINE configs[3]): n_trees searches in lockstep, fixed playouts per move, leaf batch 256.

    python tools/bench_mcts.py [--trees 64] [--playouts 16384] [--leaf-batch 256] [--moves 1] [--no-cache]

One JSON line: playouts/s over all trees (CUDA events around iago_mcts_search), per-wave time, and what the
value cache saved.  Root = the opening after move 19, colour 2 to move.
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def multi_stream(args):
    """args.streams engines, each with trees/streams trees on its own CUDA stream, driven from host threads: the select kernel of
    one group overlaps the trunk / rollout launches of the others."""
    import threading, time
    import torch
    import iago_b200
    from iago_b200.search import SearchPool
    mdir = os.path.join(ROOT, "baseline", "_ref", "models")
    T = args.trees // args.streams
    p1, p2 = (1 << 19) | (1 << 27) | (1 << 28) | (1 << 35), 1 << 36
    kw = dict(slot_policy=0, slot_value=1, lmbda=0.5, c_puct=1, n_thr=15, leaf_batch=args.leaf_batch, virtual_loss=1.0,
              precision=args.precision, cache_value=not args.no_cache, seed=1)
    groups = []
    for i in range(args.streams):
        eng = iago_b200.Engine(0)
        eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
        eng.load_net(1, os.path.join(mdir, "value_model.npz"))
        eng.load_rollout_npz(os.path.join(mdir, "rollout_model.npz"))
        groups.append((eng, SearchPool(T, max_nodes=args.max_nodes, max_leaf_batch=args.leaf_batch, tree_id0=i * T, engine=eng),
                       torch.cuda.Stream()))

    def work(eng, pool, stream, n):
        with torch.cuda.stream(stream):
            pool.set_roots(p1, p2, 2, reset_tree=True)
            pool.search(n, **kw)

    for n in (2 * args.leaf_batch, args.playouts):     # warm-up, then the timed move
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(e, p, s, n)) for e, p, s in groups]
        [t.start() for t in th]
        [t.join() for t in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    best = [int(p.root_stats()[2][0]) for _, p, _ in groups]
    print(json.dumps({"metric": "mcts_playouts_per_s", "value": T * args.streams * args.playouts / dt, "trees": T * args.streams,
                      "streams": args.streams, "playouts_per_move": args.playouts, "leaf_batch": args.leaf_batch,
                      "ms_per_move": dt * 1e3, "cache_value": not args.no_cache, "best": best, "timing": "host wall clock around the threads"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trees", type=int, default=64)
    ap.add_argument("--playouts", type=int, default=16384)
    ap.add_argument("--leaf-batch", type=int, default=256)
    ap.add_argument("--moves", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--precision", type=int, default=3)
    ap.add_argument("--no-cache", action="store_true")
    ap.add_argument("--max-nodes", type=int, default=65536)
    ap.add_argument("--streams", type=int, default=1, help="independent engines (contexts + streams) searching concurrently from host threads")
    args = ap.parse_args()
    import torch
    import iago_b200
    from iago_b200 import boards
    from iago_b200.search import SearchPool
    if args.streams > 1:
        return multi_stream(args)
    eng = iago_b200.Engine(0)
    mdir = os.path.join(ROOT, "baseline", "_ref", "models")
    eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
    eng.load_net(1, os.path.join(mdir, "value_model.npz"))
    eng.load_rollout_npz(os.path.join(mdir, "rollout_model.npz"))
    pool = SearchPool(args.trees, max_nodes=args.max_nodes, max_leaf_batch=args.leaf_batch, engine=eng)
    # the opening after colour 1 plays 19 (known answer in SURVEY.md §4)
    p1 = (1 << 19) | (1 << 27) | (1 << 28) | (1 << 35)
    p2 = 1 << 36
    kw = dict(slot_policy=0, slot_value=1, lmbda=0.5, c_puct=1, n_thr=15, leaf_batch=args.leaf_batch, virtual_loss=1.0,
              precision=args.precision, cache_value=not args.no_cache, seed=1)
    res = []
    for rep in range(args.warm + 1):
        pool.set_roots(p1, p2, 2, reset_tree=True)
        ms = []
        for mv in range(args.moves):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            pool.search(args.playouts, **kw)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            visits, q, best = pool.root_stats()
            if mv + 1 < args.moves:
                pool.advance(best)
        res = ms
    waves = -(-args.playouts // args.leaf_batch)
    tot = args.trees * args.playouts * args.moves
    t = sum(res) / 1e3
    nodes = [len(pool.export_tree(i)["n"]) for i in range(min(args.trees, 4))]
    print(json.dumps({"metric": "mcts_playouts_per_s", "value": tot / t, "trees": args.trees, "playouts_per_move": args.playouts,
                      "leaf_batch": args.leaf_batch, "moves": args.moves, "ms_per_move": res, "ms_per_wave": res[0] / waves,
                      "cache_value": not args.no_cache, "precision": args.precision, "overflows": pool.overflows(),
                      "nodes_in_first_trees": nodes, "root_visits_tree0": visits[0][visits[0] > 0].tolist(), "best": best[:4].tolist()}))


if __name__ == "__main__":
    main()
