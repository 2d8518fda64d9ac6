"""Small invocations of the integer kernels for compute-sanitizer (run on the GPU box):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py rollout mcts selfplay
    compute-sanitizer --tool racecheck python tools/sanitize_run.py rollout mcts

Sizes are tiny (the tools slow a kernel down 10-100x); results are still checked against the CPU oracle where one call does it.
profiles/run_sanitize.sh is the recipe whose logs are committed under profiles/.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import iago_b200
from iago_b200 import Rng, boards
from iago_b200.search import SearchPool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
what = sys.argv[1:] or ["rollout", "mcts"]
eng = iago_b200.Engine(0)
z = np.load(os.path.join(ROOT, "tests", "golden", "rollout_model.npz"))
eng.load_rollout(z["conv1/W"], z["bias2/b"])
dev = torch.device("cuda", 0)
mdir = os.path.join(ROOT, "baseline", "_ref", "models")

if "rollout" in what:
    from oracle import cref
    n = 1024
    p1 = np.full(n, boards.START_P1, np.uint64)
    p2 = np.full(n, boards.START_P2, np.uint64)
    out = eng.rollout_host(p1, p2, np.ones(n, np.uint8), rng=Rng.philox(seed=7), want_moves=True)
    st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
    ref = cref.simulate_batch(st, 1, z["conv1/W"], z["bias2/b"], mode=cref.RNG_PHILOX, seed=7, threads=0)
    assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all()
    # FORCED replay of the same games (the rules-only kernel variant) and the uniform-replay variant
    d1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev)
    d2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
    col = torch.ones(n, dtype=torch.uint8, device=dev)
    forced = torch.from_numpy(out["moves"]).to(dev)
    r2 = eng.rollout(d1, d2, col, rng=Rng.replay_moves(forced))
    assert (r2["final_p1"].cpu().numpy().view(np.uint64) == out["final_p1"]).all()
    u = torch.rand(n, 64, dtype=torch.float64, device=dev)
    eng.rollout(d1, d2, col, rng=Rng.replay_uniforms(u))
    torch.cuda.synchronize()
    print(f"rollout: {n} games x 3 rng modes ok", flush=True)

if "mcts" in what:
    eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
    eng.load_net(1, os.path.join(mdir, "value_model.npz"))
    p1, p2 = (1 << 19) | (1 << 27) | (1 << 28) | (1 << 35), 1 << 36
    for trees, batch, playouts in ((2, 16, 96), (1, 64, 256), (2, 1, 40)):
        pool = SearchPool(trees, max_nodes=4096, max_leaf_batch=batch, engine=eng)
        pool.set_roots(p1, p2, 2)
        pool.search(playouts, slot_policy=0, slot_value=1, leaf_batch=batch, seed=3)
        visits, _, best = pool.root_stats()
        assert int(visits.sum()) > 0
        pool.advance(best)
        pool.search(batch * 2, slot_policy=0, slot_value=1, leaf_batch=batch, seed=4)
        torch.cuda.synchronize()
        print(f"mcts: {trees} trees, leaf batch {batch}, {playouts} playouts + re-root ok -> {best.tolist()}", flush=True)
        pool.close()

if "selfplay" in what:
    eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
    res = eng.selfplay(0, 0, 64, greedy=False, rng=Rng.philox(seed=5, stream_id=1))
    torch.cuda.synchronize()
    print("selfplay: 64 sampled games ok,", res["stats"], flush=True)

if "reinforce" in what:
    # K6: REINFORCE and value gradients on the tensor-core path (fused backward chain, weight-gradient kernel, head kernels, Adam)
    from iago_b200 import network
    from iago_b200.train_rl import ReinforceTrainer
    from iago_b200.train_value import ValueTrainer
    opp_net = network.SLPolicy().load(os.path.join(mdir, "RL", "model0.npz"))
    tr = ReinforceTrainer(os.path.join(mdir, "rl_model.npz"), max_positions=2048)
    d = tr.play_set(opp_net, 32, seed=1)
    tr.gradient(d["own"], d["opp"], d["action"], d["reward"])
    tr.update()
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy()
    assert np.isfinite(g).all() and np.abs(g).max() > 0
    print(f"reinforce: 32 games, {d['own'].numel()} positions, gradient + Adam ok", flush=True)
    m = 150   # not a multiple of the weight-gradient slice count: ragged slices
    vt = ValueTrainer(os.path.join(mdir, "value_model.npz"), max_positions=256, slot=7, seed=11)
    y = torch.sign(torch.randn(m, device=dev))
    vt.gradient(d["own"][:m].contiguous(), d["opp"][:m].contiguous(), y)
    torch.cuda.synchronize()
    gv = vt.grad.cpu().numpy()
    assert np.isfinite(gv).all() and np.abs(gv).max() > 0
    print(f"value: {m} positions, gradient ok", flush=True)
