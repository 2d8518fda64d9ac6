"""The CPU oracle (oracle/othello_ref.c) pinned against vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py): rules, perft, Philox known answers, rollout policy, full Simulate trajectories."""
import numpy as np
import pytest


def mask_of(actions):
    m = 0
    for a in actions:
        m |= 1 << int(a)
    return m


def test_legal_actions_match_reference(cref, golden_rules):
    g = golden_rules
    for s, c, m in zip(g["state"], g["color"], g["legal_mask"]):
        assert mask_of(cref.legal_actions(s.astype(np.float32), int(c))) == int(m)


def test_known_answers(cref, golden_rules):
    # SURVEY.md §4 known answers, also stored by the generator from the reference's own rules
    assert cref.legal_actions(cref.start_board(), 1) == [19, 26, 37, 44] == golden_rules["start_legal_1"].tolist()
    assert cref.legal_actions(cref.start_board(), 2) == [20, 29, 34, 43] == golden_rules["start_legal_2"].tolist()
    s = cref.start_board()
    cref.place_stone(s, 19, 1)
    expect = np.zeros((8, 8), np.float32)
    expect[2, 3] = expect[3, 3] = expect[3, 4] = expect[4, 3] = 1
    expect[4, 4] = 2
    assert (s == expect).all()


def test_place_stone_matches_reference(cref, golden_rules):
    g = golden_rules
    for s, c, a, after in zip(g["ps_state"], g["ps_color"], g["ps_action"], g["ps_after"]):
        t = s.astype(np.float32).reshape(8, 8).copy()
        cref.place_stone(t, int(a), int(c))
        assert (t.reshape(64).astype(np.uint8) == after).all()
    s = cref.start_board()
    cref.place_stone(s, -1, 1)  # pass is a no-op (game.py:181-182)
    assert (s == cref.start_board()).all()


def test_perft(cref, golden_rules):
    got = [cref.perft(cref.start_board(), 1, d) for d in range(1, 9)]
    assert got[:6] == golden_rules["perft"].tolist()            # reference rules, depth 1..6
    assert got == [4, 12, 56, 244, 1396, 8200, 55092, 390216]   # the published Othello perft series


def test_philox_known_answers(cref):
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert cref.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert cref.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert cref.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    # the engine's stream: draw d of a game = word d & 3 of the block with counter (game_lo, game_hi, d >> 2, stream); u = word / 2^32
    words = cref.philox([5, 0, 7 >> 2, 0], [11, 0])
    assert cref.philox_uniform(11, 5, 7, 0) == words[7 & 3] / 4294967296.0


def test_exp32_accuracy(cref):
    xs = np.concatenate([-np.logspace(-6, np.log10(80), 4000), [0.0, -80.0, -80.5, -1000.0]]).astype(np.float32)
    got = np.array([cref.exp32_neg(x) for x in xs], np.float64)
    ref = np.exp(xs.astype(np.float64))
    ok = xs >= -80
    rel = np.abs(got[ok] - ref[ok]) / ref[ok]
    assert rel.max() < 2.5e-7          # ~2 ulp_f32
    assert (got[~ok] == 0).all()


def test_rollout_policy_matches_reference(cref, golden_nets, rollout_weights):
    W, b = rollout_weights
    g = golden_nets
    worst = 0.0
    for s, c, p in zip(g["state"], g["color"], g["rollout_prob"]):
        logits = cref.rollout_logits(s.astype(np.float32), int(c), W, b).astype(np.float64)
        e = np.exp(logits - logits.max())
        worst = max(worst, np.abs(e / e.sum() - p).max())
    assert worst < 1e-6
    # known answer at the opening
    l = cref.rollout_logits(cref.start_board(), 1, W, b).astype(np.float64)
    p = np.exp(l - l.max()); p /= p.sum()
    assert np.allclose(p[[19, 26, 37, 44]], [0.1242045, 0.1243336, 0.1263941, 0.1267480], atol=2e-7)


def test_simulate_trajectories_match_reference(cref, golden_simulate, rollout_weights):
    """1,200 unmodified-reference Simulate games replayed from the uniforms their np.random seed yields."""
    W, b = rollout_weights
    g = golden_simulate
    r = cref.simulate_batch(g["start"].astype(np.float32), g["color"].astype(np.int32), W, b,
                            mode=cref.RNG_UNIFORMS, uniforms=g["uniforms"], threads=0)
    assert (r["moves"] == g["moves"]).all()
    assert (r["n_moves"] == g["n_moves"]).all()
    assert (r["results"] == g["result"]).all()
    assert (r["final"].reshape(-1, 64).astype(np.uint8) == g["final"]).all()


def test_simulate_forced_replay(cref, golden_simulate, rollout_weights):
    W, b = rollout_weights
    g = golden_simulate
    r = cref.simulate_batch(g["start"][:200].astype(np.float32), g["color"][:200].astype(np.int32), W, b,
                            mode=cref.RNG_FORCED, forced=g["moves"][:200], threads=2)
    assert (r["final"].reshape(-1, 64).astype(np.uint8) == g["final"][:200]).all()
    assert (r["results"] == g["result"][:200]).all()


def test_philox_stream_is_thread_count_invariant(cref, rollout_weights):
    W, b = rollout_weights
    st = np.tile(cref.start_board().reshape(1, 64), (512, 1))
    a = cref.simulate_batch(st, 1, W, b, seed=99, game_id0=1000, threads=1)
    c = cref.simulate_batch(st, 1, W, b, seed=99, game_id0=1000, threads=4)
    assert (a["moves"] == c["moves"]).all() and (a["results"] == c["results"]).all()
    # sharding invariance: games [256, 512) computed alone equal the same ids inside the big batch
    d = cref.simulate_batch(st[256:], 1, W, b, seed=99, game_id0=1256, threads=2)
    assert (d["moves"] == a["moves"][256:]).all()


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_live_reference_agrees_with_golden_file(golden_simulate):
    """Re-runs a few games through the reference itself to show the committed fixture is what it produces."""
    import subprocess, sys, os, json
    code = r'''
import sys, json, numpy as np
sys.path.insert(0, "ORACLE")
import ref_harness
m = ref_harness.load()
out = []
for seed in (12345, 12346, 12347):
    np.random.seed(seed)
    s = np.zeros([8, 8], np.float32); s[4, 3] = s[3, 4] = 1; s[3, 3] = s[4, 4] = 2
    sim = m["mcts_self_play"].Simulate(s)
    out.append([int(sim(1)), sim.state.astype(int).reshape(64).tolist()])
print(json.dumps(out))
'''.replace("ORACLE", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    g = golden_simulate
    for i, (r, final) in enumerate(got):
        assert g["seed"][i] == 12345 + i and r == g["result"][i] and final == g["final"][i].tolist()


def test_canon_exp_and_sampler_choice(cref, rollout_weights):
    """canon_exp (the table builder shared, operation for operation, by the oracle and the library) is exp to ~1 ulp of a double,
    and the committed rollout weights select the table sampler; absurd weights select the exp32 fallback."""
    xs = np.concatenate([np.linspace(-300, 300, 2001), [0.0, -1e-9, 1e-9, 88.7, -87.3]])
    got = np.array([cref.canon_exp(x) for x in xs])
    assert np.max(np.abs(got - np.exp(xs)) / np.exp(xs)) < 5e-16
    W, b = rollout_weights
    assert cref.policy_is_fast(W, b)
    assert not cref.policy_is_fast(W * 40, b) and not cref.policy_is_fast(W * np.float32("nan"), b)


def test_two_sided_cdf_is_numpy_choice(cref, rollout_weights):
    """The canonical sampling rule (running sums over the legal cells 0..31 ascending and 63..32 descending, see the header of
    oracle/othello_ref.c) is, in exact arithmetic, np.random.choice's searchsorted(cdf, u, 'right') over the masked, renormalised
    softmax (mcts_self_play.py:100-110).  Checked against a float64 numpy restatement of that line on random boards, built from the
    fp32 logits: the picks agree everywhere except where u sits within 2e-6 of a cdf edge (rounding the logit sum to fp32, which the
    canonical weights E0 * E1 * EB do not do, moves an edge by up to ~1e-7 * |logit|; probes 1e-5 either side of the edges are included)."""
    W, b = rollout_weights
    rng = np.random.default_rng(5)
    checked = near = 0
    for trial in range(4000):
        fill = rng.random()
        r = rng.random(64)
        st = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0)).astype(np.float32)
        color = int(rng.integers(1, 3))
        acts = cref.legal_actions(st, color)
        if len(acts) < 2:
            continue
        logits = np.asarray(cref.rollout_logits(st, color, W, b), np.float64).reshape(64)
        p = np.exp(logits - logits.max())
        mask = np.zeros(64)
        mask[acts] = 1.0
        p = p * mask
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        us = np.concatenate([rng.random(6), [0.0, 1 - 2.0**-53], cdf[acts[:-1]][:3] + 1e-5, cdf[acts[:-1]][:3] - 1e-5])
        for u in us:
            if not 0.0 <= u < 1.0:
                continue
            want = int(np.searchsorted(cdf, u, side="right"))
            got = cref.rollout_sample(st, color, W, b, float(u))
            if np.min(np.abs(cdf[acts] - u)) < 2e-6:
                near += 1
                continue
            checked += 1
            assert got == want, (trial, u, got, want)
    assert checked > 30000 and near < checked // 5   # (edges of cells with probability below 1e-5 crowd each other)


def test_torch_nets_equal_numpy_nets(golden_nets):
    """oracle/nets_torch.py (used for the >= 50,000-position GPU comparison) against oracle/nets.py, which is pinned to the reference's
    own outputs: same graph, fp32 conv summation order differs (<= 1e-4 on logits that span 0..130), fp64 agrees to 1e-4 as well."""
    import os
    import torch
    from conftest import MODELS
    from oracle import nets, nets_torch
    if not os.path.isfile(os.path.join(MODELS, "sl_model.npz")):
        pytest.skip("baseline/_ref/models not present")
    x = golden_nets["x"][:200].astype(np.float32)
    p = nets.load_params(os.path.join(MODELS, "sl_model.npz"))
    ref = nets.sl_logits(p, x)
    assert np.abs(nets_torch.sl_logits(p, x) - ref).max() <= 1e-4
    assert np.abs(nets_torch.sl_logits(p, x, torch.float64) - ref).max() <= 1e-4
    pv = nets.load_params(os.path.join(MODELS, "value_model.npz"))
    assert np.abs(nets_torch.value(pv, x) - nets.value(pv, x)).max() <= 1e-5
    # and against the reference's own forward (golden file)
    e = np.exp(nets_torch.sl_logits(p, x).astype(np.float64)); e /= e.sum(axis=1, keepdims=True)
    assert np.abs(e - golden_nets["sl_prob"][:200]).max() <= 1e-5


def big_uniforms(g):
    """The uniforms the 21,000 games of simulate_big.npz consumed: game i ran under np.random.seed(seed0 + i) (oracle/gen_golden_big.py)."""
    seed0 = int(g["seed0"])
    return np.stack([np.random.RandomState(seed0 + i).random_sample(64) for i in range(len(g["moves"]))])


def test_21000_reference_games_replayed(cref, rollout_weights):
    """simulate_big.npz: 21,000 full games of the unmodified mcts_self_play.Simulate (14,000 from the opening, 7,000 from mid-game positions
    with either side to move) replayed by the C oracle from the np.random uniforms of each game's seed: every move, result and final
    board identical.  This is the pin of the sampling rule (the two-sided cdf) on 1.08 M reference plies."""
    from conftest import load_golden
    from iago_b200 import boards
    W, b = rollout_weights
    g = load_golden("simulate_big")
    n = len(g["moves"])
    assert n >= 20000
    start = boards.from_bitboards(g["start_p1"], g["start_p2"]).reshape(n, 64)
    r = cref.simulate_batch(start.astype(np.float32), g["color"].astype(np.int32), W, b, mode=cref.RNG_UNIFORMS, uniforms=big_uniforms(g), threads=0)
    assert (r["n_moves"] == g["n_moves"]).all()
    assert (r["moves"][:, :60] == g["moves"]).all()
    assert (r["results"] == g["result"]).all()
    f1, f2 = cref.to_bitboards(r["final"])
    assert (f1 == g["final_p1"]).all() and (f2 == g["final_p2"]).all()
    # how close the recorded draws come to a cdf edge is what makes the double running sum necessary: report the count of passes / early ends
    print(f"{n} reference games, {int(g['n_moves'].sum())} plies; {(g['n_moves'] < 60 - (np.unpackbits(g['start_p1'].view(np.uint8)).reshape(n, 64).sum(1) + np.unpackbits(g['start_p2'].view(np.uint8)).reshape(n, 64).sum(1) - 4)).sum()} games ended with empty cells")
