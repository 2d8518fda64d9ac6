#!/usr/bin/env python
"""bench.py — rollout plies/s of the lockstep rollout kernel (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--games G]

One "step" = one pass of the hot path over one batch: 65,536 rollout-policy games per GPU from the opening,
colour 1 first, models/rollout_model.npz, Philox uniforms keyed by the global game id (weak scaling: every
rank plays its own 65,536 games per step; no data-path collective — only the timing/plies all-reduce).

Printed JSON (one line, rank 0):
  value      plies/s over all ranks, inputs resident in HBM, per-step CUDA events on the launching stream,
             L2 flushed between steps, max over ranks
  e2e        the same metric through the host-buffer C-ABI calls on pinned caller buffers, transfers inside the timed region:
             value = batches streamed with iago_rollout_host_submit / _wait (three in flight), synchronous_call = one blocking
             iago_rollout_host per step
  roofline   achieved int32 lane-ops/s (676 per ply) vs the integer-issue peak measured live with
             iago_measure_int_peak; the path is issue-bound, not HBM-bound (roofline_hbm shows why)
  cpu_baseline   the CPU oracle port (oracle/othello_ref.c, pthreads over all cores) on a bounded sample,
             plus the unmodified Python reference under the chainer stand-in when baseline/_ref is present
  --impl reference   times the CPU port alone (the reference is pure Python and cannot be "compiled")
  selfplay / mcts / reinforce / valuegen   the other BASELINE configs as extra sections of the same line (games/s, playouts/s,
             records/s, each with its roofline fraction and the reference's CPU figure); the nets run in the inference precision 2
             (fp16 + FP8 cross terms), `precision3` inside selfplay / mcts is the same workload in the parity precision;
             `selfplay.sampled` is the batch with sampled moves and switched openings (games that differ)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMES_PER_STEP = 65536
OPS_PER_PLY = 676          # int32 ALU lane-ops per ply: movegen 314 + flip 352 + ~10 (SURVEY.md §8d, step-by-step 6-step flood)
BYTES_PER_GAME = 17 + 21   # p1,p2,colour in; final p1,p2,n_moves,result out


def ncu_traffic(key="rollout_pair_kernel"):
    """Per-launch figures of the headline kernel from the committed ncu capture (profiles/ncu_traffic.json, written by
    profiles/ncu_traffic.py with the capture's date and command): dram__bytes_read + write, and the warp instructions the ALU pipe
    executed (`key` = rollout_pair_kernel for the Philox launch, rollout_pair_kernel_forced for the rules-only replay).  None when the
    file is missing or is for another launch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]
        return t if t.get("games_per_launch") == GAMES_PER_STEP else None
    except Exception:
        return None


def alu_lane_ops_per_ply(key, plies_per_launch):
    """ALU-pipe lane-operations per ply the kernel really issued in the committed capture (warp instructions x 32 / plies of one
    launch of the same workload); None without a capture."""
    t = ncu_traffic(key)
    if not t or not t.get("alu_pipe_warp_inst_per_launch") or not plies_per_launch:
        return None
    return 32.0 * t["alu_pipe_warp_inst_per_launch"] / plies_per_launch

METRIC = "rollout_plies_per_s"


def rollout_weights():
    for p in (os.path.join(ROOT, "baseline", "_ref", "models", "rollout_model.npz"),
              os.path.join(ROOT, "tests", "golden", "rollout_model.npz")):
        if os.path.isfile(p):
            z = np.load(p)
            return z["conv1/W"], z["bias2/b"]
    raise FileNotFoundError("rollout_model.npz")


class ClockSampler:
    """NVML sampling of SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self.ok = [], set(), None, threading.Event(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            # a few closely spaced samples for the short regions (the headline step is ~10 ms), then 20 Hz (NVML queries take a driver
            # lock; the long sections read device counters from the host every pair of turns)
            time.sleep(0.002 if len(self.samples) < 8 else 0.05)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": statistics.median(self.samples), "sm_min_mhz": min(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_port_throughput(n_games, threads=0, seed=777):
    """plies/s of the CPU oracle port on `n_games` games of the bench workload. Returns (plies/s, threads, plies, s)."""
    from oracle import cref
    cref.build()
    W, b = rollout_weights()
    st = np.tile(cref.start_board().reshape(1, 64), (n_games, 1))
    t0 = time.perf_counter()
    r = cref.simulate_batch(st, 1, W, b, mode=cref.RNG_PHILOX, seed=seed, game_id0=0, want_moves=False, threads=threads)
    dt = time.perf_counter() - t0
    plies = int(r["n_moves"].sum())
    return plies / dt, int(r["threads"]), plies, dt


def python_reference_throughput(n_games=12, procs=1):
    """The UNMODIFIED reference Simulate (baseline/_ref copies) under the chainer stand-in, one process, one core."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "mcts_self_play.py")):
        return None
    code = f"""
import os, sys, time, json
os.environ["IAGO_REFERENCE"] = {ref!r}
sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
import numpy as np, ref_harness
m = ref_harness.load()
Sim = m["mcts_self_play"].Simulate
plies = 0
np.random.seed(1)
t0 = time.perf_counter()
for g in range({n_games}):
    s = np.zeros([8, 8], np.float32); s[4, 3] = s[3, 4] = 1; s[3, 3] = s[4, 4] = 2
    sim = Sim(s); sim(1)
    plies += int((sim.state != 0).sum()) - 4
dt = time.perf_counter() - t0
print(json.dumps({{"plies_per_s": plies / dt, "games": {n_games}, "seconds": dt}}))
"""
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    try:
        if procs <= 1:
            res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
            return json.loads(res.stdout.strip().splitlines()[-1])
        # the same script fanned out over all host cores, one process per core (the reference itself is single-process);
        # rate = plies of all processes / the longest process's own timed region (interpreter start-up excluded, as above)
        ps = [subprocess.Popen([sys.executable, "-c", code.replace("np.random.seed(1)", f"np.random.seed({1 + i})")],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for i in range(procs)]
        outs = [json.loads(p_.communicate(timeout=600)[0].strip().splitlines()[-1]) for p_ in ps]
        secs = max(o["seconds"] for o in outs)
        plies = sum(o["plies_per_s"] * o["seconds"] for o in outs)
        return {"plies_per_s": plies / secs, "games": n_games * procs, "seconds": secs, "cores": procs}
    except Exception as e:
        return {"error": repr(e)[:200]}


def _run_reference_snippet(body, timeout=600):
    """Runs `body` (python source that prints one JSON line) with the UNMODIFIED reference modules loaded as `m`."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "mcts_self_play.py")):
        return None
    code = f"""
import os, sys, time, json
os.environ["IAGO_REFERENCE"] = {ref!r}
sys.path.insert(0, {os.path.join(ROOT, 'oracle')!r})
import numpy as np, ref_harness
m = ref_harness.load()
""" + body
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    try:
        res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, env=env)
        return json.loads(res.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"error": repr(e)[:200]}


def python_reference_selfplay(n_games=3):
    """The unmodified src/rl_self_play.Game (sl_model vs sl_model) under the chainer stand-in, one core."""
    return _run_reference_snippet(f"""
net, ser = m["network"], m["chainer"].serializers
a = net.SLPolicy(); ser.load_npz("./models/sl_model.npz", a)
b = net.SLPolicy(); ser.load_npz("./models/sl_model.npz", b)
np.random.seed(3)
t0 = time.perf_counter()
for g in range({n_games}):
    m["rl_self_play"].Game(a, b)()
dt = time.perf_counter() - t0
print(json.dumps({{"games_per_s": {n_games} / dt, "games": {n_games}, "seconds": dt}}))
""")


def python_reference_mcts(n_playouts=60):
    """The unmodified MCTS.playout (sl/value/rollout nets, lmbda 0.5) from the position after move 19, one core."""
    return _run_reference_snippet(f"""
gf = m["game"].GameFunctions
s = np.zeros([8, 8], np.float32); s[4, 3] = s[3, 4] = 1; s[3, 3] = s[4, 4] = 2
gf.place_stone(s, 19, 1)
mc = m["MCTS"].MCTS()
np.random.seed(5)
t0 = time.perf_counter()
for k in range({n_playouts}):
    mc.playout(s.copy(), 2, mc.root)
dt = time.perf_counter() - t0
print(json.dumps({{"playouts_per_s": {n_playouts} / dt, "playouts": {n_playouts}, "seconds": dt}}))
""")


def cpu_baseline_block(budget_s=12.0):
    rate, threads, _, _ = cpu_port_throughput(8192)
    n = int(min(max(rate * budget_s / 60.0, 8192), 4_000_000))
    rate, threads, plies, dt = cpu_port_throughput(n)
    out = {"value": rate, "unit": "plies/s", "cores": threads, "kind": "port",
           "sample": f"{n} games ({plies} plies) of the bench workload in {dt:.1f} s, oracle/othello_ref.c, {threads} pthreads"}
    py = python_reference_throughput()
    if py:
        out["python_reference"] = dict(py, cores=1, note="unmodified reference mcts_self_play.Simulate under the numpy "
                                       "chainer stand-in (Chainer itself is not installable)")
        cores = os.cpu_count() or 1
        if cores > 1 and "error" not in py:
            out["python_reference_all_cores"] = dict(python_reference_throughput(12, procs=cores),
                                                     note="the same, one process per host core running side by side")
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = args.games or GAMES_PER_STEP   # the same 65,536-game step as the GPU arm (about 0.3 s of CPU work per step on 16 threads)
    for _ in range(args.warmup):
        cpu_port_throughput(min(n, 2048))
    plies = 0
    t = 0.0
    threads = 0
    for i in range(args.steps):
        r, threads, p, dt = cpu_port_throughput(n, seed=1000 + i)
        plies += p
        t += dt
    v = plies / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "plies/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 bitboards + f32 policy", "data": "synthetic",
            "config": {"workload": "65,536 lockstep rollout-policy games per GPU per step from the opening, colour 1 first, "
                                   "rollout_model.npz, Philox4x32-10 uniforms keyed by global game id (BASELINE configs[1])",
                       "games_per_step_per_gpu": n, "implementation": "oracle/othello_ref.c on all host threads (the reference itself is "
                       "single-process Python: 4.9e3 plies/s, see cpu_baseline.python_reference of the GPU arm)"},
            "cpu_baseline": {"value": v, "unit": "plies/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x {n} games, oracle/othello_ref.c, {threads} pthreads"},
            "e2e": {"value": v, "unit": "plies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import iago_b200
    from iago_b200 import Rng, boards

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    eng = iago_b200.Engine(local_rank)
    eng.load_rollout(*rollout_weights())
    n = args.games or GAMES_PER_STEP

    p1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev)
    p2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
    col = torch.ones(n, dtype=torch.uint8, device=dev)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    out = dict(result=torch.empty(n, dtype=torch.int8, device=dev), final_p1=torch.empty_like(p1),
               final_p2=torch.empty_like(p2), n_moves=torch.empty(n, dtype=torch.int32, device=dev), moves=None)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(i, timed=None):
        gid0 = (i * world + rank) * n  # = parallel.game_id0: global game ids, results do not depend on the number of ranks
        flush.fill_(i & 0xFF)
        if timed is not None:
            timed[0].record()
        eng.rollout(p1, p2, col, rng=Rng.philox(seed=args.seed, game_id0=gid0), counters=counters, out=out)
        if timed is not None:
            timed[1].record()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    counters.zero_()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            step(args.warmup + i, evs[i])
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    t_dev = sum(step_ms) / 1e3
    plies = int(counters[0].item())
    turns = int(counters[1].item())

    # end-to-end through the host-buffer C-ABI call
    # (caller buffers in pinned host memory, as the bench contract asks; the pageable-buffer figure is reported beside it)
    def e2e_run(pinned):
        hp1, hp2, hcol, hout = eng.rollout_host_buffers(n, pinned=pinned)
        hp1[:], hp2[:], hcol[:] = boards.START_P1, boards.START_P2, 1
        for i in range(max(args.warmup, 3)):
            eng.rollout_host(hp1, hp2, hcol, rng=Rng.philox(seed=args.seed + 1, game_id0=(i * world + rank) * n), out=hout)
        barrier()
        done = 0
        t0 = time.perf_counter()
        for i in range(args.steps):
            eng.rollout_host(hp1, hp2, hcol, rng=Rng.philox(seed=args.seed + 1, game_id0=((100 + i) * world + rank) * n), out=hout)
            done += int(hout["counters"][0])
        t = time.perf_counter() - t0
        barrier()
        return done, t

    # the same batches streamed: iago_rollout_host_submit / _wait with three batches in flight, each with its own pinned buffers —
    # every step still copies its 17 B per game in and its 21 B per game out, but on the copy engines, beside another step's kernel
    def e2e_streamed(lanes=3):
        bufs = [eng.rollout_host_buffers(n) for _ in range(lanes)]
        for hp1, hp2, hcol, _ in bufs:
            hp1[:], hp2[:], hcol[:] = boards.START_P1, boards.START_P2, 1

        def run(steps, base):
            done = 0
            for i in range(steps):
                ln = i % lanes
                if i >= lanes:
                    done += int(eng.rollout_host_wait(ln)["counters"][0])
                hp1, hp2, hcol, hout = bufs[ln]
                eng.rollout_host_submit(ln, hp1, hp2, hcol, rng=Rng.philox(seed=args.seed + 1, game_id0=((base + i) * world + rank) * n), out=hout)
            for i in range(max(0, steps - lanes), steps):
                done += int(eng.rollout_host_wait(i % lanes)["counters"][0])
            return done

        run(max(args.warmup, 3), 200)
        barrier()
        t0 = time.perf_counter()
        done = run(args.steps, 300)
        t = time.perf_counter() - t0
        barrier()
        return done, t

    pageable_plies, t_pageable = e2e_run(False)
    sync_plies, t_sync = e2e_run(True)
    e2e_plies, t_e2e = e2e_streamed()

    # movegen / flip alone: the same kernel replaying the games' own move logs (no policy, no sampling) — the integer path
    # whose issue utilisation the north star asks for separately
    log = eng.rollout(p1, p2, col, rng=Rng.philox(seed=args.seed, game_id0=rank * n), want_moves=True)
    forced = Rng.replay_moves(log["moves"])
    mg_counters = torch.zeros(2, dtype=torch.int64, device=dev)
    for i in range(3):
        eng.rollout(p1, p2, col, rng=forced, out=out)
    barrier()
    mg_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in mg_ev:
        flush.fill_(1)
        a.record()
        eng.rollout(p1, p2, col, rng=forced, counters=mg_counters, out=out)
        b.record()
    barrier()
    t_mg = sum(a.elapsed_time(b) for a, b in mg_ev) / 1e3
    mg_plies = int(mg_counters[0].item())

    ops_impl = alu_lane_ops_per_ply("rollout_pair_kernel", plies / args.steps) if n == GAMES_PER_STEP else None   # this rank's launches
    ops_impl_mg = alu_lane_ops_per_ply("rollout_pair_kernel_forced", mg_plies / args.steps) if n == GAMES_PER_STEP else None

    if dist is not None:
        tt = torch.tensor([t_dev, t_e2e, t_wall], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, t_wall = tt.tolist()
        cc = torch.tensor([plies, turns, e2e_plies], dtype=torch.int64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        plies, turns, e2e_plies = cc.tolist()

    if rank == 0:
        int_peak = eng.measure_int_peak(4096)
        kernel_s = statistics.mean(step_ms) / 1e3
        plies_per_launch = plies / (args.steps * world)
        achieved = OPS_PER_PLY * plies_per_launch / kernel_s
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = BYTES_PER_GAME * n / kernel_s / 1e9
        line = {
            "metric": METRIC, "value": plies / t_dev, "unit": "plies/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64 bitboards + f32 policy", "data": "synthetic",
            "config": {"workload": "65,536 lockstep rollout-policy games per GPU per step from the opening, colour 1 first, "
                                   "rollout_model.npz, Philox4x32-10 uniforms keyed by global game id (BASELINE configs[1])",
                       "games_per_step_per_gpu": n, "l2": "flushed between steps (256 MiB fill); inputs 1.1 MB",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks"},
            "games_per_s": (args.steps * world * n) / t_dev,
            "plies_per_game": plies / (args.steps * world * n), "turns_per_game": turns / (args.steps * world * n),
            "wall_s_timed_region": t_wall,
            "e2e": {"value": e2e_plies / t_e2e, "unit": "plies/s", "h2d_bytes_per_step": 17 * n,
                    "d2h_bytes_per_step": 21 * n + 16,
                    "api": "iago_rollout_host_submit / _wait on pinned caller buffers, three batches in flight: per step H2D copies "
                           "(17 B per game), the kernel, D2H copies (21 B per game) on the step's own stream; wall clock from the "
                           "first submit to the last wait",
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "note": "can exceed `value`: `value` times every launch alone between two L2 flushes (its own ramp and tail of one "
                            "147-CTA wave), streamed launches on different streams overlap those with the neighbouring batch",
                    "synchronous_call": {"value": sync_plies / t_sync, "ms_per_step": 1e3 * t_sync / args.steps,
                                         "api": "iago_rollout_host, one blocking call per step on pinned, device-mapped buffers: one "
                                                "launch whose loads / stores cross PCIe + stream sync", "scope": "rank 0"},
                    "pageable_buffers": {"value": pageable_plies / t_pageable, "ms_per_step": 1e3 * t_pageable / args.steps,
                                         "api": "same call on pageable numpy arrays: packed into pinned staging, 4-chunk "
                                                "H2D / kernel / D2H pipeline", "scope": "rank 0"}},
            "gpu_launches": args.steps * world,
            "clocks": clk.summary(),
            "roofline": {"bound": "alu", "kernel": "rollout_pair_kernel<PHILOX> (two lanes per game)", "achieved": achieved / 1e12,
                         "peak": int_peak / 1e12, "unit": "Tint32op/s", "frac": achieved / int_peak,
                         "traffic": (ncu_traffic() or {}).get("dram_bytes_per_launch") if n == GAMES_PER_STEP else None,
                         "traffic_source": ncu_traffic(),
                         "alu_lane_ops_per_ply_issued": ops_impl,
                         "frac_alu_pipe_issued": (ops_impl * plies_per_launch / kernel_s / int_peak) if ops_impl else None,
                         "note": "issue-bound path: 676 algorithmic int32 lane-ops/ply (SURVEY 8d: 6-step flood formulation) x plies per launch / mean launch "
                                 "time; peak = SHF+LOP3 micro-kernel measured in this run (iago_measure_int_peak); traffic = "
                                 "dram__bytes_read+write per launch read from profiles/ncu_traffic.json (the committed ncu --set full capture; "
                                 "algorithmic bytes per launch: 38 B x 65,536 games = 2.49 MB; outputs stay in L2 during the capture); "
                                 "alu_lane_ops_per_ply_issued = ALU-pipe warp instructions of that capture x 32 / plies (rules + policy + "
                                 "sampling), frac_alu_pipe_issued = what they occupy of the ALU pipe at this run's speed"},
            "roofline_movegen": {"bound": "alu", "kernel": "rollout_pair_kernel<FORCED> (legal_moves + flips + pass/terminal/score only, moves "
                                 "replayed from a 64 B/game log)", "plies_per_s": mg_plies / t_mg,
                                 "achieved": OPS_PER_PLY * mg_plies / t_mg / 1e12, "peak": int_peak / 1e12, "unit": "Tint32op/s",
                                 "frac": OPS_PER_PLY * mg_plies / t_mg / int_peak, "traffic": None, "scope": "rank 0",
                                 "alu_lane_ops_per_ply_issued": ops_impl_mg,
                                 "frac_alu_pipe_issued": (ops_impl_mg * mg_plies / t_mg / int_peak) if ops_impl_mg else None,
                                 "note": "the kernel finds flips and east moves by carry propagation and floods the other directions in parallel-prefix "
                                         "form, which takes about half the ALU operations of the step-by-step form SURVEY 8d counts (676), so the "
                                         "fraction on the 676 count exceeds 1; alu_lane_ops_per_ply_issued = ALU-pipe warp instructions of the "
                                         "committed ncu capture of this launch x 32 / plies, frac_alu_pipe_issued = what they occupy of the ALU pipe "
                                         "at this run's speed (the capture itself: sm__inst_executed_pipe_alu in profiles/ncu_traffic.json)"},
            "roofline_hbm": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                             "frac": hbm_ach / hbm_peak, "traffic": None,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                             "note": "38 algorithmic bytes per game; shown to document that HBM is not the limiter"},
        }
    extra = {}
    for name, fn in (("selfplay", section_selfplay), ("mcts", section_mcts), ("reinforce", section_reinforce),
                     ("valuegen", section_valuegen)):
        if name in args.sections:
            try:
                extra[name] = fn(eng, args, rank, world, dev, dist, barrier)
            except Exception as e:  # a failed extra section must not lose the headline line
                extra[name] = {"error": repr(e)[:300]}
    if rank == 0:
        # the all-core CPU leg runs last, after every GPU section has been timed
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_block()
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


SL_FLOP, VALUE_FLOP = 122847232, 122994944   # per position


def model_path(name):
    p = os.path.join(ROOT, "baseline", "_ref", "models", name)
    if not os.path.isfile(p):
        raise FileNotFoundError(f"{p}: run oracle/fetch_ref.py in the build container (baseline/_ref ships with gpurun)")
    return p


def bf16_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained"
    except Exception:
        return 1386.3, "fallback (B200_PROFILING.md sustained bf16)"


def section_selfplay(eng, args, rank, world, dev, dist, barrier):
    """BASELINE configs[2]: SL-policy greedy self-play, 16,384-game lockstep batch per GPU, sl_model.npz both sides — plus the same batch
    with SAMPLED moves and the head/tail-switched openings of src/train_rl.py:43-46 (the greedy games are all the same game)."""
    import torch
    from iago_b200 import Rng, boards
    from iago_b200.train_rl import SWITCH_CELLS
    n = args.selfplay_games
    eng.load_net(0, model_path("sl_model.npz"))
    reps = max(1, args.selfplay_steps)
    peak, src = bf16_peak()

    def timed(greedy, precision, init=None, steps=reps):
        """`steps` self-play steps; returns (seconds max over ranks, games, positions, last result dict, per-step ms, clocks)."""
        kw = dict(greedy=greedy, precision=precision)
        if init is not None:
            kw.update(init_p1=init[0], init_p2=init[1])
        res = None
        for w in range(3):   # warm-up at full size, holding the previous step's result like the timed loop does: both sets of result
            res = eng.selfplay(0, 0, n, rng=Rng.philox(seed=1 + w, stream_id=1), **kw)   # tensors exist before the timing starts (with
            # one discarded warm-up step the second timed step paid a cudaMalloc of the record buffers: 180-250 ms instead of 138)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        fwd = evaluated = 0
        with ClockSampler(dev.index or 0) as clk:
            for i in range(steps):
                ev[i][0].record()
                res = eng.selfplay(0, 0, n, rng=Rng.philox(seed=args.seed, game_id0=(i * world + rank) * n, stream_id=1), **kw)
                ev[i][1].record()
                fwd += res["stats"]["forwards"]
                evaluated += res["stats"]["positions"]
            barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        tt = torch.tensor([sum(step_ms) / 1e3], dtype=torch.float64, device=dev)
        cc = torch.tensor([steps * n, evaluated], dtype=torch.int64, device=dev)   # positions the nets really evaluated (request lists)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        return float(tt[0]), int(cc[0]), int(cc[1]), res, step_ms, clk.summary(), fwd / steps

    def wdl(res):
        w = torch.stack([(res["result"] == 1).sum(), (res["result"] == 0).sum(), (res["result"] == -1).sum()]).to(torch.int64)
        if dist is not None:
            dist.all_reduce(w, op=dist.ReduceOp.SUM)   # win statistics: the only cross-GPU exchange of this workload
        return w.tolist()

    # (a) the headline: greedy, inference precision (2)
    t, games, positions, res, step_ms, clocks, fpg = timed(True, None)
    ach = positions * SL_FLOP / t / 1e12
    # (b) the same in precision 3 (the parity setting): same games?
    t3, games3, _, res3, _, _, _ = timed(True, 3, steps=1)
    same_games = bool((res3["final_p1"] == res["final_p1"]).all() and (res3["final_p2"] == res["final_p2"]).all()) if reps == 1 else None
    if same_games is None:   # compare like with like: step 0 of both runs
        r2 = eng.selfplay(0, 0, n, greedy=True, precision=None, rng=Rng.philox(seed=args.seed, game_id0=rank * n, stream_id=1))
        same_games = bool((res3["final_p1"] == r2["final_p1"]).all() and (res3["final_p2"] == r2["final_p2"]).all())
    # (c) sampled moves, odd games from a head/tail-switched opening: 16,384 different games of different lengths per GPU
    p1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev)
    p2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
    cells = torch.tensor(SWITCH_CELLS, dtype=torch.int64, device=dev)[torch.randint(0, 4, (n,), device=dev, generator=torch.Generator(device=dev).manual_seed(args.seed + rank))]
    odd = (torch.arange(n, device=dev) & 1) == 1
    p2 = torch.where(odd, p2 | (torch.ones_like(p2) << cells), p2)   # a colour-2 stone dropped without flipping (src/train_rl.py:43-46)
    ts, games_s, positions_s, res_s, step_ms_s, clocks_s, fpg_s = timed(False, None, init=(p1, p2))
    n_final = torch.tensor([len(torch.unique(torch.stack([res_s["final_p1"], res_s["final_p2"]], 1), dim=0))], dtype=torch.int64, device=dev)
    ach_s = positions_s * SL_FLOP / ts / 1e12
    out = {"metric": "selfplay_games_per_s", "value": games / t, "unit": "games/s", "ms_per_step": 1e3 * t / reps,
           "config": {"workload": "SL-policy greedy self-play, lockstep batch per GPU, sl_model.npz vs sl_model.npz "
                                  "(BASELINE configs[2])", "games_per_step_per_gpu": n, "steps": reps,
                      "precision": "2 (inference default): fp16 main product + FP8 cross terms, 2 MMA units per K step; logits within 1e-2 "
                                   "(measured 3e-3), legal arg-max identical (tests/test_nets_gpu.py)"},
           "positions_per_s": positions / t, "trunk_forwards_per_game": fpg,
           "step_ms_rank0": [round(x, 2) for x in step_ms], "clocks": clocks,
           "precision3": {"value": games3 / t3, "unit": "games/s", "ms_per_step": 1e3 * t3,
                          "note": "the same step with the nets in precision 3 (fp16 hi/lo split, 3 MMAs per K step, ~1e-4 on the logits): the parity setting",
                          "same_final_boards_as_default_precision": same_games},
           "last_step_w_d_l": wdl(res),
           "roofline": {"bound": "tensor", "kernel": "trunk_kernel", "achieved": ach, "peak": peak * world, "unit": "TFLOP/s",
                        "frac": ach / (peak * world), "traffic": None, "peak_source": src,
                        "note": "algorithmic FLOP (122,847,232 per position, counted once) over the whole self-play step incl. the turn "
                                "kernels; precision 2 issues 2x this many MMA FLOP-equivalents"},
           "sampled": {"value": games_s / ts, "unit": "games/s", "ms_per_step": 1e3 * ts / reps, "positions_per_s": positions_s / ts,
                       "trunk_forwards_per_game": fpg_s, "step_ms_rank0": [round(x, 2) for x in step_ms_s], "clocks": clocks_s,
                       "last_step_w_d_l": wdl(res_s), "distinct_final_boards_rank0_last_step": int(n_final[0]),
                       "roofline": {"bound": "tensor", "achieved": ach_s, "peak": peak * world, "unit": "TFLOP/s", "frac": ach_s / (peak * world)},
                       "note": "the same batch with SAMPLED moves (masked softmax, one Philox uniform per move) and every odd game started from a "
                               "head/tail-switched opening (src/train_rl.py:43-46): games differ in content and length, finished games idle "
                               "in their lanes until the longest one ends (what REINFORCE self-play does)"}}
    if rank == 0 and world == 1 and not args.no_cpu:
        py = python_reference_selfplay()
        if py:
            out["cpu_baseline"] = dict(py, unit="games/s", value=py.get("games_per_s"), cores=1, kind="reference",
                                       sample="unmodified src/rl_self_play.Game under the numpy chainer stand-in, sampled moves")
    return out


def section_mcts(eng, args, rank, world, dev, dist, barrier):
    """BASELINE configs[3]: PV-MCTS, fixed 16K playouts per move, virtual-loss leaf batch 256, many trees per GPU."""
    import torch
    from iago_b200.search import SearchPool
    eng.load_net(0, model_path("sl_model.npz"))
    eng.load_net(1, model_path("value_model.npz"))
    T, B, N = args.mcts_trees, 256, args.mcts_playouts
    pool = SearchPool(T, max_nodes=32768, max_leaf_batch=B, tree_id0=rank * T, engine=eng)
    p1, p2 = (1 << 19) | (1 << 27) | (1 << 28) | (1 << 35), 1 << 36   # the opening after colour 1 plays 19
    out = {}
    for cache, prec in ((True, None), (False, None), (True, 3)):
        kw = dict(slot_policy=0, slot_value=1, lmbda=0.5, c_puct=1, n_thr=15, leaf_batch=B, virtual_loss=1.0, precision=prec,
                  cache_value=cache, seed=args.seed)
        pool.set_roots(p1, p2, 2, reset_tree=True)
        pool.search(2 * B, **kw)   # warm-up waves
        pool.set_roots(p1, p2, 2, reset_tree=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pool.search(N, **kw)
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt[0])
        visits, q, best = pool.root_stats()
        out["precision3" if prec == 3 else "cached" if cache else "uncached"] = {
            "value": T * N * world / t, "ms_per_move": 1e3 * t, "best_move_tree0": int(best[0]), "pool_overflows": pool.overflows()}
    peak, src = bf16_peak()
    unc = out["uncached"]["value"]
    res = {"metric": "mcts_playouts_per_s", "value": out["cached"]["value"], "unit": "playouts/s",
           "config": {"workload": "PV-MCTS, sl/value/rollout nets, lmbda 0.5, c_puct 1, n_thr 15, leaf batch 256, virtual loss 1, root = "
                                  "opening after move 19, one move (BASELINE configs[3])", "trees_per_gpu": T, "playouts_per_move": N},
           "value_cache_on": out["cached"], "value_cache_off": out["uncached"],
           "precision3": dict(out["precision3"], note="the same search (value cache on) with the nets in precision 3 (fp16 hi/lo split, ~1e-4 on the "
                              "logits) instead of the inference default 2 (fp16 + FP8 cross terms, 3e-3; north-star bar 1e-2)"),
           "roofline": {"bound": "tensor", "kernel": "trunk_kernel (value net, cache off: one forward per playout)",
                        "achieved": unc * VALUE_FLOP / 1e12, "peak": peak * world, "unit": "TFLOP/s",
                        "frac": unc * VALUE_FLOP / 1e12 / (peak * world), "traffic": None, "peak_source": src,
                        "note": "algorithmic value-net FLOP per playout x playouts/s with the value cache OFF (the reference's work per "
                                "playout); with the cache on most playouts need no net evaluation and the search is bound by the select "
                                "kernel and the rollouts"}}
    if rank == 0 and world == 1 and not args.no_cpu:
        py = python_reference_mcts()
        if py:
            res["cpu_baseline"] = dict(py, unit="playouts/s", value=py.get("playouts_per_s"), cores=1, kind="reference",
                                       sample="unmodified MCTS.playout under the numpy chainer stand-in")
    pool.close()
    # One game searched by ALL ranks at once (the optional exchange for a single tree): every rank runs N / world playouts
    # on the same root under its own global tree id (its rollouts draw from different Philox streams), the root visit counts [65] are
    # summed over the ranks (parallel.root_parallel_moves) and the move is chosen from the sum.  At one rank this is the single-tree
    # latency figure below; the ratio to it is what root parallelism buys.
    if world > 1:
        from iago_b200 import parallel
        share = N // world
        rp = SearchPool(1, max_nodes=65536, max_leaf_batch=B, tree_id0=20_000_000 + rank, engine=eng)
        kw = dict(slot_policy=0, slot_value=1, lmbda=0.5, c_puct=1, n_thr=15, leaf_batch=B, virtual_loss=1.0, precision=None,
                  cache_value=True, seed=args.seed)
        rp.set_roots(p1, p2, 2, reset_tree=True)
        rp.search(2 * B, **kw)
        rp.set_roots(p1, p2, 2, reset_tree=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rp.search(share, **kw)
        v, _, _ = rp.root_stats()
        vs, best_rp = parallel.root_parallel_moves(torch.from_numpy(v).to(dev))
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res["root_parallel"] = {"ms_per_move": 1e3 * float(tt[0]), "playouts_per_rank": share, "playouts_total": share * world,
                                "best_move": int(best_rp[0]), "root_visits_total": int(vs[0].sum()),
                                "note": "one tree per rank on the same root, root visit counts all-reduced (one int64[65] sum), move from the sum"}
        rp.close()
    if rank == 0:
        # one tree (the reference's usage: one game, one search per move): latency of a 16,384-playout move
        one = SearchPool(1, max_nodes=65536, max_leaf_batch=B, tree_id0=10_000_000, engine=eng)
        kw = dict(slot_policy=0, slot_value=1, lmbda=0.5, c_puct=1, n_thr=15, leaf_batch=B, virtual_loss=1.0, precision=None,
                  cache_value=True, seed=args.seed)
        one.set_roots(p1, p2, 2, reset_tree=True)
        one.search(2 * B, **kw)
        one.set_roots(p1, p2, 2, reset_tree=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one.search(N, **kw)
        b.record()
        torch.cuda.synchronize()
        t1 = a.elapsed_time(b) / 1e3
        res["single_tree"] = {"playouts_per_s": N / t1, "ms_per_move": 1e3 * t1, "best_move": int(one.root_stats()[2][0]),
                              "note": "one search tree on one GPU (latency-bound: 64 dependent waves of 256 leaves)"}
        one.close()
        # the same move budget as 8 root-parallel trees on this GPU (MCTS(root_trees=8)): 8 dependent waves instead of 64
        R = 8
        rp1 = SearchPool(R, max_nodes=65536, max_leaf_batch=B, tree_id0=30_000_000, engine=eng)
        rp1.set_roots(p1, p2, 2, reset_tree=True)
        rp1.search(2 * B, **kw)
        rp1.set_roots(p1, p2, 2, reset_tree=True)
        a.record()
        rp1.search(N // R, **kw)
        v8 = rp1.root_stats()[0].astype(np.int64).sum(axis=0)
        b.record()
        torch.cuda.synchronize()
        t8 = a.elapsed_time(b) / 1e3
        res["single_game_root_trees_8"] = {"playouts_per_s": N / t8, "ms_per_move": 1e3 * t8, "best_move": int(np.argmax(v8[:64])),
                                           "note": "one game, 8 independent trees of 2,048 playouts on one GPU, move from the summed root visits "
                                                   "(opt-in MCTS(root_trees=8); not the reference's single-tree algorithm)"}
        rp1.close()
    return res


def section_reinforce(eng, args, rank, world, dev, dist, barrier):
    """BASELINE configs[4]: train_rl.py-style REINFORCE self-play, rl_model.npz vs RL/model0.npz, games sharded over the ranks,
    [gradient | loss | count] all-reduced over NCCL once per update through the library (iago_comm_allreduce_sum_f32 enqueued on the
    stream; the Adam step reads the reduced count on the device).  Two measurements: `updates` back-to-back 2,048-game updates, and
    the config at its full size — 1,048,576 games on 8 GPUs = 131,072 games per GPU (weak scaling: every rank plays its share)."""
    import torch
    from iago_b200 import network, parallel
    from iago_b200.train_rl import ReinforceTrainer
    n = args.reinforce_games
    opp = network.SLPolicy(device=eng.device).load(model_path("RL/model0.npz"))
    comm = parallel.Communicator.from_process_group(eng) if world > 1 else None
    tr = ReinforceTrainer(model_path("rl_model.npz"), alpha=1e-3, max_positions=8192, device=eng.device, comm=comm)
    tr.train_set(opp, n_games=min(n, 256), seed=args.seed, game_id0=parallel.game_id0(0, rank, world, n))   # warm-up update
    tr.train_set(opp, n_games=n, seed=args.seed, game_id0=parallel.game_id0(1, rank, world, n))             # and one at full size (workspaces)
    barrier()

    def run(steps, first_step):
        stats = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            stats = tr.train_set(opp, n_games=n, seed=args.seed, game_id0=parallel.game_id0(first_step + i, rank, world, n),
                                 want_stats=(i == steps - 1))
        b.record()
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / 1e3], dtype=torch.float64, device=dev)
        parallel.all_reduce_max_(tt)
        return float(tt[0]), stats

    steps = max(1, args.reinforce_steps)
    t, last = run(steps, 2)
    out = {"metric": "reinforce_games_per_s", "value": steps * n * world / t, "unit": "games/s", "ms_per_update": 1e3 * t / steps,
           "config": {"workload": "REINFORCE sets: self-play (sampled, odd games head/tail switched) + gradient + all-reduce + Adam/WD, "
                                  "rl_model.npz learner vs RL/model0.npz (BASELINE configs[4])", "games_per_update_per_gpu": n,
                      "updates": steps, "gradient_arithmetic": "tcgen05: fp16 hi/lo forward, fused bf16 hi/lo data-gradient chain, fp16 (scaled, hi/lo dY) weight-gradient GEMMs; fp32 accumulate",
                      "collective": "ncclAllReduce of 960,770 floats enqueued through iago_comm_allreduce_sum_f32; count read on the device" if comm else "none (one rank)"},
           "positions_last_update": last.get("positions"), "last": last, "allreduce_bytes_per_update": (960768 + 2) * 4}
    if args.reinforce_1m_games > 0:
        per_gpu = args.reinforce_1m_games // 8          # the config is 1 M games on 8 GPUs: every rank plays an eighth of it
        big_steps = max(1, per_gpu // n)
        t1, last1 = run(big_steps, 2 + steps)
        out["config4_full_size"] = {"games": big_steps * n * world, "games_per_gpu": big_steps * n, "updates": big_steps, "seconds": t1,
                                    "value": big_steps * n * world / t1, "unit": "games/s", "ms_per_update": 1e3 * t1 / big_steps, "last": last1,
                                    "note": "BASELINE configs[4] at its per-GPU size (1,048,576 games / 8 GPUs): at --gpus 8 this is the full 1 M games"}
    if comm is not None:
        comm.close()
    tr.close()
    return out


def section_valuegen(eng, args, rank, world, dev, dist, barrier):
    """value-data generation (value_self_play.SelfPlay x gen_value_data.py), lockstep batch per GPU."""
    import numpy as np
    import torch
    from iago_b200 import Rng
    from iago_b200.engine import STREAM_VALUEGEN
    n = args.valuegen_games
    eng.load_net(0, model_path("sl_model.npz"))
    eng.load_net(2, model_path("rl_model.npz"))
    stop = torch.from_numpy(np.random.RandomState(args.seed + rank).randint(4, 64, size=n).astype(np.int32)).to(dev)
    eng.value_selfplay(0, 2, stop[:2048].contiguous(), rng=Rng.philox(seed=1, stream_id=STREAM_VALUEGEN))   # warm-up
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    res = eng.value_selfplay(0, 2, stop, rng=Rng.philox(seed=args.seed, game_id0=rank * n, stream_id=STREAM_VALUEGEN))
    b.record()
    barrier()
    tt = torch.tensor([a.elapsed_time(b) / 1e3], dtype=torch.float64, device=dev)
    usable = (res["rec_action"] >= 0).sum().to(torch.int64).reshape(1)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(usable, op=dist.ReduceOp.SUM)
    t = float(tt[0])
    out = {"metric": "value_records_per_s", "value": n * world / t, "unit": "games/s", "ms_per_step": 1e3 * t,
           "config": {"workload": "value_self_play.SelfPlay(stop_num ~ randint(4, 64)): sl_model.npz to stop_num, one random move, "
                                  "rl_model.npz to the end (gen_value_data.py)", "games_per_step_per_gpu": n},
           "records_with_a_move": int(usable[0]), "turns": res["stats"]["turns"], "trunk_launches": res["stats"]["forwards"]}
    if rank == 0 and world == 1 and not args.no_cpu:
        import time
        from oracle import nets, valuegen_ref
        psl, prl = nets.load_params(model_path("sl_model.npz")), nets.load_params(model_path("rl_model.npz"))
        f = lambda p: (lambda st, c: nets.sl_logits(p, nets.planes_from_state(st[None], c))[0])
        rs = np.random.RandomState(1)
        t0 = time.perf_counter()
        k = 3
        for g in range(k):
            valuegen_ref.play(int(rs.randint(4, 64)), f(psl), f(prl), rs.random_sample(200))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": k / dt, "unit": "games/s", "cores": 1, "kind": "port",
                               "sample": f"{k} games of oracle/valuegen_ref.py with the fp32 numpy nets (the reference file itself is dead at "
                                         "HEAD: it imports a deleted module and its softmax overflows on sl_model.npz)"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=0, help="games per step per GPU (default 65,536 for both arms)")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--reinforce-games", type=int, default=2048)
    ap.add_argument("--reinforce-steps", type=int, default=8)
    ap.add_argument("--reinforce-1m-games", type=int, default=1048576, help="total games of the configs[4] run on 8 GPUs (each rank plays an eighth); 0 = skip")
    ap.add_argument("--valuegen-games", type=int, default=16384)
    ap.add_argument("--sections", default="rollout,selfplay,mcts,reinforce,valuegen", help="extra sections to run after the headline rollout bench")
    ap.add_argument("--selfplay-games", type=int, default=16384)
    ap.add_argument("--selfplay-steps", type=int, default=4)
    ap.add_argument("--mcts-trees", type=int, default=256)
    ap.add_argument("--mcts-playouts", type=int, default=16384)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun exactly as the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
