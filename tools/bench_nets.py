#!/usr/bin/env python
"""Throughput of the fused trunk kernel (SLPolicy / Value forward) — BASELINE configs[2] batch shape.

    python tools/bench_nets.py [--n 16384] [--steps 10] [--precision 3] [--kind policy|value]

Prints one JSON line: positions/s, algorithmic TFLOP/s (122,847,232 FLOP/position SL, 122,994,944 value),
MMA TFLOP/s actually issued (x3 for the hi/lo split), fraction of the measured bf16 peak (MEASURED_PEAKS.json).
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", type=int, default=3)
    ap.add_argument("--kind", default="policy")
    args = ap.parse_args()
    import torch
    import iago_b200
    eng = iago_b200.Engine(0)
    mdir = os.path.join(ROOT, "baseline", "_ref", "models")
    eng.load_net(0, os.path.join(mdir, "sl_model.npz" if args.kind == "policy" else "value_model.npz"))
    g = np.load(os.path.join(ROOT, "tests", "golden", "rules.npz"))
    from iago_b200 import boards
    p1, p2 = boards.to_bitboards(g["state"][:4096])
    idx = np.arange(args.n) % len(p1)
    dev = torch.device("cuda", 0)
    d1 = torch.from_numpy(p1[idx].view(np.int64)).to(dev)
    d2 = torch.from_numpy(p2[idx].view(np.int64)).to(dev)
    col = torch.from_numpy(g["color"][:4096][idx].astype(np.uint8)).to(dev)
    out = torch.empty((args.n, 64) if args.kind == "policy" else (args.n,), dtype=torch.float32, device=dev)
    fwd = (lambda: eng.policy_forward(0, d1, d2, col, probs=False, precision=args.precision, out=out)) if args.kind == "policy" \
        else (lambda: eng.value_forward(0, d1, d2, col, precision=args.precision, out=out))
    for _ in range(args.warmup):
        fwd()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        a.record(); fwd(); b.record()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs]
    t = float(np.mean(ms)) / 1e3
    flop = 122847232 if args.kind == "policy" else 122994944
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("bf16_tflops", 1590.0)
    alg = flop * args.n / t / 1e12
    print(json.dumps({"kernel": "trunk_kernel", "kind": args.kind, "n": args.n, "precision": args.precision,
                      "ms": t * 1e3, "ms_min": min(ms), "positions_per_s": args.n / t, "algorithmic_tflops": alg,
                      "mma_tflops_issued": alg * (3 if args.precision == 3 else 1),
                      "frac_of_measured_bf16_peak_algorithmic": alg / peak,
                      "frac_of_measured_bf16_peak_issued": alg * (3 if args.precision == 3 else 1) / peak, "peak_tflops": peak}))


if __name__ == "__main__":
    main()
