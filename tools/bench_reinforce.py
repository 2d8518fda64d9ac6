#!/usr/bin/env python
"""Throughput of the REINFORCE gradient (K6) and of one train_rl-style set (self-play + gradient + Adam).

    python tools/bench_reinforce.py [--positions 8192] [--games 1024]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=8192)
    ap.add_argument("--games", type=int, default=1024)
    args = ap.parse_args()
    import torch
    import iago_b200
    from iago_b200 import network
    from iago_b200.train_rl import ReinforceTrainer
    mdir = os.path.join(ROOT, "baseline", "_ref", "models")
    opp = network.SLPolicy().load(os.path.join(mdir, "RL", "model0.npz"))
    tr = ReinforceTrainer(os.path.join(mdir, "rl_model.npz"), max_positions=args.positions)
    d = tr.play_set(opp, args.games, seed=1)
    m = d["own"].numel()
    reps = -(-args.positions // m)
    own, oppb = d["own"].repeat(reps)[:args.positions].contiguous(), d["opp"].repeat(reps)[:args.positions].contiguous()
    act, rew = d["action"].repeat(reps)[:args.positions].contiguous(), d["reward"].repeat(reps)[:args.positions].contiguous()
    tr.gradient(own, oppb, act, rew)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    tr.gradient(own, oppb, act, rew)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    flop = 3 * 122847232 * args.positions
    t0 = time.perf_counter()
    st = tr.train_set(opp, n_games=args.games, seed=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"gradient_ms": ms, "positions": args.positions, "positions_per_s": args.positions / ms * 1e3,
                      "fp32_tflops": flop / ms / 1e9, "train_set": dict(st, games=args.games, seconds=dt, games_per_s=args.games / dt)}))


if __name__ == "__main__":
    main()
