"""Trunk throughput by precision mode (run on the GPU box): python tools/bench_nets_prec.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import iago_b200
from iago_b200 import Rng, boards
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
eng = iago_b200.Engine(0)
mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "models")
eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
eng.load_net(1, os.path.join(mdir, "value_model.npz"))
z = np.load(os.path.join(mdir, "rollout_model.npz")); eng.load_rollout(z["conv1/W"], z["bias2/b"])
dev = torch.device("cuda", 0)
p1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev); p2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
col = torch.ones(n, dtype=torch.uint8, device=dev)
forced = torch.randint(0, 64, (n, 20), dtype=torch.int8, device=dev)   # scatter the positions: 20 arbitrary placements each
out = eng.rollout(p1, p2, col, rng=Rng.replay_moves(forced))
q1, q2 = out["final_p1"], out["final_p2"]
for prec in [int(x) for x in os.environ.get("PRECS", "3,2,1").split(",")]:
    for slot, name, fn in ((0, "policy", eng.policy_forward), (1, "value", eng.value_forward)):
        fn(slot, q1, q2, col, precision=prec)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn(slot, q1, q2, col, precision=prec)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"precision {prec} {name:6s}: {ms:.3f} ms per {n} positions = {n / ms * 1e3:.4g} positions/s, {n * 122.85e6 / ms / 1e9:.0f} algorithmic TFLOP/s", flush=True)
