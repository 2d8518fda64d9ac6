"""iago_rollout_host end to end (run on the GPU box): pageable caller buffers (staged, 4-chunk pipeline) vs page-locked ones
(used in place: one launch over mapped memory for Philox games).  The chunk-count / copy-mode variants measured while this path
was designed are recorded in csrc/rollout.cu next to the code that chooses between them."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import iago_b200
from iago_b200 import Rng, boards
from bench import rollout_weights

n = 65536
eng = iago_b200.Engine(0)
eng.load_rollout(*rollout_weights())


def bufs(pinned):
    def mk(shape, dt, fill=None):
        if pinned:
            t = torch.empty(shape, dtype=dt, pin_memory=True)
            a = t.numpy()
        else:
            a = np.empty(shape, {torch.int64: np.int64, torch.uint8: np.uint8, torch.int8: np.int8, torch.int32: np.int32}[dt])
        if fill is not None:
            a[...] = fill
        return a
    p1 = mk(n, torch.int64, np.int64(np.uint64(boards.START_P1))).view(np.uint64)
    p2 = mk(n, torch.int64, np.int64(np.uint64(boards.START_P2))).view(np.uint64)
    col = mk(n, torch.uint8, 1)
    out = dict(result=mk(n, torch.int8), final_p1=mk(n, torch.int64).view(np.uint64), final_p2=mk(n, torch.int64).view(np.uint64),
               n_moves=mk(n, torch.int32), moves=None, counters=np.zeros(2, np.uint64))
    return p1, p2, col, out


def run(tag, pinned):
    p1, p2, col, out = bufs(pinned)
    for i in range(5):
        eng.rollout_host(p1, p2, col, rng=Rng.philox(seed=7, game_id0=0), out=out)
    chk = (int(out["final_p1"].sum()), int(out["n_moves"].sum()), int(out["result"].astype(np.int64).sum()))
    torch.cuda.synchronize()
    steps = 50
    t0 = time.perf_counter()
    plies = 0
    for i in range(steps):
        eng.rollout_host(p1, p2, col, rng=Rng.philox(seed=7, game_id0=(i + 1) * n), out=out)
        plies += int(out["counters"][0])
    dt = time.perf_counter() - t0
    print(f"{tag:34s} {1e3 * dt / steps:7.4f} ms/step  {plies / dt:.4g} plies/s  check {chk}", flush=True)


run("pageable buffers (staged)", False)
run("pinned buffers (in place)", True)
