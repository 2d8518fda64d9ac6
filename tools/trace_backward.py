"""Debug: per-layer hand-over times of the backward data-gradient chain (trunk_kernel<1>; needs -DIAGO_TRUNK_TRACE).
The chain is the last trunk launch of a gradient call, so its trace is what iago_debug_trace returns afterwards."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from iago_b200 import network
from iago_b200.train_rl import ReinforceTrainer
mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "models")
opp = network.SLPolicy().load(os.path.join(mdir, "RL", "model0.npz"))
tr = ReinforceTrainer(os.path.join(mdir, "rl_model.npz"), max_positions=8192)
d = tr.play_set(opp, 512, seed=3)
n = min(8192, d["own"].numel())
for _ in range(2):
    tr.gradient(d["own"][:n], d["opp"][:n], d["action"][:n], d["reward"][:n]); torch.cuda.synchronize()
import iago_b200
lib = iago_b200.load_library()
buf = np.zeros(4096, np.uint64)
lib.iago_debug_trace(C.c_void_p(buf.ctypes.data), 4096)
t = buf.reshape(-1, 8)[:4 * 9].reshape(4, 9, 8).astype(np.int64)
t0 = t[1, 0, 0]
print(f"backward chain, {n} positions: cycles relative to tile 1 layer 0 (issuer: chunk0 go, chunk1 go, last commit | epilogue: acc ready, pass0 done, pass1 done)")
for l in range(7):
    e = t[1, l] - t0
    prev = t[1, l - 1] if l else t[0, 6]
    print(f"  layer {l}: chunk0-go {e[0]:7d} chunk1-go {e[1]:7d} last-commit {e[2]:7d} | acc-ready {e[3]:7d} pass0-done {e[4]:7d} pass1-done {e[5]:7d} | "
          f"layer span {t[1, l, 3] - prev[3]:6d}  issuing {t[1, l, 2] - t[1, l, 0]:6d}  drain {t[1, l, 3] - t[1, l, 2]:5d}  prev-acc-ready -> chunk0-go {t[1, l, 0] - prev[3]:6d}")
u = buf[2048:2048 + 144].reshape(18, 8).astype(np.int64)
base = u[0, 2]
print("units of tile 1 layer 2: producer empty-ready, tma issued | issuer full-ready, commit issued")
for i in range(18):
    print(f"   unit {i:2d}: producer {u[i, 0] - base:7d} {u[i, 1] - base:7d} | issuer {u[i, 2] - base:7d} {u[i, 3] - base:7d}  (issue span {u[i, 3] - u[i, 2]:5d}; unit-to-unit {u[i, 2] - u[i - 1, 2] if i else 0:5d})")
