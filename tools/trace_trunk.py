"""Debug: per-layer hand-over times of the trunk pipeline (needs a library built with -DIAGO_TRUNK_TRACE)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import iago_b200
from iago_b200 import boards
eng = iago_b200.Engine(0)
mdir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "models")
eng.load_net(0, os.path.join(mdir, "sl_model.npz"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda", 0)
p1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev); p2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
col = torch.ones(n, dtype=torch.uint8, device=dev)
for prec in (3, 1, 2):
    eng.policy_forward(0, p1, p2, col, precision=prec); torch.cuda.synchronize()
    eng.policy_forward(0, p1, p2, col, precision=prec); torch.cuda.synchronize()
    buf = np.zeros(4096, np.uint64)
    eng.lib.iago_debug_trace(C.c_void_p(buf.ctypes.data), 4096)
    t = buf.reshape(-1, 8)[:4 * 9].reshape(4, 9, 8).astype(np.int64)
    u = buf[2048:2048 + 144].reshape(18, 8).astype(np.int64)
    base = u[0, 2]
    print(f"precision {prec}: units of tile 1 layer 2 (cycles from the issuer's first 'full'): producer empty-ready, tma issued | issuer full-ready, commit issued")
    for i in range(18):
        print(f"   unit {i:2d}: producer {u[i, 0] - base:7d} {u[i, 1] - base:7d} | issuer {u[i, 2] - base:7d} {u[i, 3] - base:7d}   (issue span {u[i, 3] - u[i, 2]:5d}; unit-to-unit {u[i, 2] - u[i - 1, 2] if i else 0:5d})")
    ee = buf[3072:3072 + 72].reshape(9, 8).astype(np.int64)
    t0 = t[1, 0, 0]
    print(f"precision {prec}: cycles relative to tile 1 layer 0 (events: chunk0 go, chunk1 go, commit issued | acc ready, pass0 done, pass1 done)")
    for tile in (1,):
        for l in range(8):
            e = t[tile, l] - t0
            prev = t[tile, l - 1] if l else t[tile - 1, 7]
            print(f"  tile {tile} layer {l}: issuer chunk0-go {e[0]:7d} chunk1-go {e[1]:7d} last-commit {e[2]:7d} | epilogue acc-ready {e[3]:7d} pass0-done {e[4]:7d} pass1-done {e[5]:7d} | "
                  f"layer span {t[tile, l, 3] - prev[3]:6d}  issuing {t[tile, l, 2] - t[tile, l, 0]:6d}  drain {t[tile, l, 3] - t[tile, l, 2]:4d}  prev-acc-ready -> this chunk0-go {t[tile, l, 0] - prev[3]:5d}")
    for l in range(1, 7):
        e = ee[l]
        print(f"  epilogue thread 0, tile 1 layer {l} pass 0: acc-ready -> start {e[0] - t[1, l, 3]:5d} | tmem loads + math {e[1] - e[0]:5d} | convert + st.shared {e[2] - e[1]:5d} | fence.proxy.async {e[3] - e[2]:5d} | syncwarp + arrive {e[4] - e[3]:5d}")
