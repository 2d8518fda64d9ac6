"""Generate tests/golden/simulate_big.npz: >= 20,000 full games of the UNMODIFIED reference mcts_self_play.Simulate.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference), about 3 min on 8 cores:

    python oracle/gen_golden_big.py [--games 20000] [--procs 8]

Why a second file: simulate.npz (1,200 games) stores the uniforms each game consumed; at 20,000 games that would be
10 MB.  A game of this file is identified by its np.random seed alone: game g runs under np.random.seed(SEED0 + g), and
the uniforms np.random.choice consumed are RandomState(SEED0 + g).random_sample(64)[:n_moves] (exactly one per stone
placed, none per pass, mcts_self_play.py:106 — asserted below for every game), so the tests regenerate them.
Stored per game: move list (int8[60]), n_moves, result, final bitboards.  Instrumentation is by wrapping
Simulate.place_stone to log the action; what the reference computes is untouched.

The mid-game third of the file starts from positions of the first games (ply 6..45), alternating the side to move, so the
sampling rule is pinned on boards MCTS leaves look like as well as on the opening.
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED0 = 5_000_000


def start_board():
    s = np.zeros([8, 8], dtype=np.float32)
    s[4, 3] = s[3, 4] = 1
    s[3, 3] = s[4, 4] = 2
    return s


def bitboards(state):
    f = np.asarray(state).reshape(64)
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    return np.uint64((w * (f == 1)).sum()), np.uint64((w * (f == 2)).sum())


def worker(job):
    lo, hi, n_open = job
    sys.path.insert(0, HERE)
    import ref_harness
    mods = ref_harness.load()
    Sim = mods["mcts_self_play"].Simulate
    gf = mods["game"].GameFunctions
    if not getattr(Sim, "_iago_logged", False):   # a pool process runs many jobs: wrap once
        orig_place = Sim.place_stone

        def place(self, state, action, color):
            self._moves.append(int(action))
            self._movers.append(int(color))
            return orig_place(self, state, action, color)

        Sim.place_stone = place
        Sim._iago_logged = True

    def run(state, color, seed):
        np.random.seed(seed)
        sim = Sim(state)
        sim._moves, sim._movers = [], []
        r = sim(color)
        nxt = np.random.random_sample()
        u = np.random.RandomState(seed).random_sample(64)
        assert nxt == u[len(sim._moves)], "np.random.choice consumed an unexpected number of draws"
        return sim, int(r)

    rows = []
    for g in range(lo, hi):
        if g < n_open:
            state, color = start_board(), 1
        else:
            # mid-game start: replay the first `ply` stones of opening game (g - n_open) — regenerated here from its seed
            src, _ = run(start_board(), 1, SEED0 + (g - n_open))
            ply = 6 + (g * 7) % 40
            state, color = start_board(), 1
            for a, who in list(zip(src._moves, src._movers))[:ply]:
                gf.place_stone(state, a, who)
                color = 3 - who
            if g % 3 == 2:
                color = 3 - color   # the "wrong" side to move: immediate passes
        s1, s2 = bitboards(state)
        sim, r = run(state.copy(), color, SEED0 + g)
        mv = np.full(60, -1, np.int8)
        mv[:len(sim._moves)] = sim._moves
        f1, f2 = bitboards(sim.state)
        rows.append((g, s1, s2, color, mv, len(sim._moves), r, f1, f2))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=20000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden", "simulate_big.npz"))
    args = ap.parse_args()
    out_path = os.path.abspath(args.out)   # the workers chdir into the reference tree
    n = args.games
    n_open = (2 * n) // 3
    chunk = 100
    jobs = [(lo, min(lo + chunk, n), n_open) for lo in range(0, n, chunk)]
    with mp.Pool(args.procs) as pool:
        rows = [r for part in pool.imap_unordered(worker, jobs) for r in part]
    rows.sort(key=lambda r: r[0])
    assert [r[0] for r in rows] == list(range(n))
    np.savez_compressed(
        out_path, seed0=np.int64(SEED0), n_open=np.int64(n_open),
        start_p1=np.array([r[1] for r in rows], np.uint64), start_p2=np.array([r[2] for r in rows], np.uint64),
        color=np.array([r[3] for r in rows], np.int8), moves=np.array([r[4] for r in rows], np.int8),
        n_moves=np.array([r[5] for r in rows], np.int8), result=np.array([r[6] for r in rows], np.int8),
        final_p1=np.array([r[7] for r in rows], np.uint64), final_p2=np.array([r[8] for r in rows], np.uint64))
    print("wrote", out_path, n, "games,", int(sum(r[5] for r in rows)), "plies", os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
