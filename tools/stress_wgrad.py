"""Stress: the gradient of one batch, repeated — every repetition must be bit-identical (fixed-order reductions; a race in the
weight-gradient producers' hand-over shows up as run-to-run differences: this is how the parity-aliasing bug of the 3-CTA cluster
experiment, DESIGN.md K6, was found)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iago_b200 import network
from iago_b200.train_rl import ReinforceTrainer
mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "models")
opp = network.SLPolicy().load(os.path.join(mdir, "RL", "model0.npz"))
tr = ReinforceTrainer(os.path.join(mdir, "rl_model.npz"), max_positions=8192)
d = tr.play_set(opp, 512, seed=3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ref = None
for n in (8192, 5000, 333, 147, 50, 3):
    n = min(n, d["own"].numel())
    for r in range(reps):
        tr.gradient(d["own"][:n], d["opp"][:n], d["action"][:n], d["reward"][:n])
        torch.cuda.synchronize()
        g = tr.grad.clone()
        if r == 0:
            ref = g
            print(f"n = {n}: |g| max {float(g.abs().max()):.6e} sum {float(g.double().sum()):.12e}")
        else:
            assert torch.equal(g, ref), f"repetition {r} differs at n = {n}: max diff {float((g - ref).abs().max())}"
print("bit-identical over", reps, "repetitions per size")
