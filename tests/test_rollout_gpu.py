"""The lockstep rollout kernel (through the C ABI) vs the reference goldens and the CPU oracle — bit exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bb(states):
    from iago_b200 import boards
    return boards.to_bitboards(states)


def test_golden_trajectories_uniform_replay(engine, golden_simulate):
    """1,200 games of the UNMODIFIED reference, replayed on the GPU from the same uniform stream."""
    from iago_b200 import Rng, boards
    g = golden_simulate
    p1, p2 = bb(g["start"])
    out = engine.rollout_host(p1, p2, g["color"].astype(np.uint8), rng=Rng.replay_uniforms(g["uniforms"]), want_moves=True)
    assert (out["moves"] == g["moves"]).all()
    assert (out["n_moves"] == g["n_moves"]).all()
    assert (out["result"] == g["result"]).all()
    final = boards.from_bitboards(out["final_p1"], out["final_p2"]).reshape(-1, 64).astype(np.uint8)
    assert (final == g["final"]).all()
    assert int(out["counters"][0]) == int(g["n_moves"].sum())


def test_21000_reference_games_uniform_replay(engine):
    """simulate_big.npz: 21,000 full games of the UNMODIFIED reference Simulate (opening and mid-game starts, both colours to move),
    replayed on the GPU from the np.random uniforms of each game's seed: 1.08 M plies, every move / result / final board identical."""
    from conftest import load_golden
    from iago_b200 import Rng
    g = load_golden("simulate_big")
    seed0, n = int(g["seed0"]), len(g["moves"])
    assert n >= 20000
    u = np.stack([np.random.RandomState(seed0 + i).random_sample(64) for i in range(n)])
    out = engine.rollout_host(g["start_p1"], g["start_p2"], g["color"].astype(np.uint8), rng=Rng.replay_uniforms(u), want_moves=True)
    assert (out["n_moves"] == g["n_moves"]).all()
    assert (out["moves"][:, :60] == g["moves"]).all()
    assert (out["result"] == g["result"]).all()
    assert (out["final_p1"] == g["final_p1"]).all() and (out["final_p2"] == g["final_p2"]).all()
    assert int(out["counters"][0]) == int(g["n_moves"].astype(np.int64).sum())


def test_golden_forced_replay(engine, golden_simulate):
    from iago_b200 import Rng, boards
    g = golden_simulate
    p1, p2 = bb(g["start"])
    out = engine.rollout_host(p1, p2, g["color"].astype(np.uint8), rng=Rng.replay_moves(g["moves"]), want_moves=True)
    final = boards.from_bitboards(out["final_p1"], out["final_p2"]).reshape(-1, 64).astype(np.uint8)
    assert (final == g["final"]).all() and (out["result"] == g["result"]).all()
    assert (out["moves"] == g["moves"]).all()


def test_65536_games_philox_vs_oracle(engine, cref, rollout_weights):
    """BASELINE config 2: 65,536 lockstep games from the opening, every trajectory bit-exact vs the CPU oracle."""
    from iago_b200 import Rng, boards
    W, b = rollout_weights
    n = 65536
    p1 = np.full(n, boards.START_P1, np.uint64)
    p2 = np.full(n, boards.START_P2, np.uint64)
    out = engine.rollout_host(p1, p2, np.ones(n, np.uint8), rng=Rng.philox(seed=2026, game_id0=0), want_moves=True)
    st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
    ref = cref.simulate_batch(st, 1, W, b, mode=cref.RNG_PHILOX, seed=2026, game_id0=0, threads=0)
    assert (out["moves"] == ref["moves"]).all()
    assert (out["result"] == ref["results"]).all()
    assert (out["n_moves"] == ref["n_moves"]).all()
    r1, r2 = bb(ref["final"])
    assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all()
    assert int(out["counters"][0]) == int(ref["n_moves"].sum()) and int(out["counters"][1]) == int(ref["n_turns"].sum())
    # shape of the workload (BASELINE.md §2 probe: 59.8 stones / game)
    assert 59.0 < out["n_moves"].mean() + 0.0 < 60.0


def test_midgame_both_colours_vs_oracle(engine, cref, rollout_weights, golden_rules):
    """Rollouts from harvested mid-game positions (what MCTS leaves look like), either side to move, incl. passes."""
    from iago_b200 import Rng
    W, b = rollout_weights
    st = golden_rules["state"][:6000].astype(np.float32)
    col = golden_rules["color"][:6000].astype(np.uint8)
    p1, p2 = bb(st)
    out = engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=5, game_id0=10**12, stream_id=3), want_moves=True)
    ref = cref.simulate_batch(st, col.astype(np.int32), W, b, mode=cref.RNG_PHILOX, seed=5, game_id0=10**12, stream=3, threads=0)
    assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all()
    r1, r2 = bb(ref["final"])
    assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all()


def test_ragged_batches_and_mixed_endgames_vs_oracle(engine, cref, rollout_weights, golden_rules):
    """The paired kernel's control flow is warp-uniform: games that have ended (or never existed: a batch that does not fill its
    last warp) run on with an empty move.  Batches of 1 ... 1,000 games that mix openings, positions a few stones from the end, full
    boards and boards without a legal move for either side, all three rng modes, every trajectory against the CPU oracle."""
    from iago_b200 import Rng, boards
    W, b = rollout_weights
    rs = np.random.RandomState(77)
    st_all = golden_rules["state"].astype(np.float32)
    filled = (st_all != 0).sum(axis=1)
    late = st_all[filled >= 56][:300]                      # a handful of empties left: games of very different lengths in one warp
    mid = st_all[(filled > 20) & (filled < 40)][:300]
    special = np.stack([np.ones(64, np.float32), np.full(64, 2, np.float32), np.zeros(64, np.float32),
                        np.concatenate([np.ones(32), np.zeros(32)]).astype(np.float32)])   # full, full, empty, no move for anyone
    pool = np.concatenate([late, mid, special, np.tile(boards.start_state().reshape(1, 64), (50, 1)).astype(np.float32)])
    for n in (1, 2, 3, 15, 16, 17, 31, 33, 257, 1000):
        st = pool[rs.randint(0, len(pool), size=n)]
        col = rs.randint(1, 3, size=n).astype(np.uint8)
        p1, p2 = bb(st)
        out = engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=n, game_id0=5 * n), want_moves=True)
        ref = cref.simulate_batch(st, col.astype(np.int32), W, b, mode=cref.RNG_PHILOX, seed=n, game_id0=5 * n, threads=0)
        assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all(), n
        assert (out["n_moves"] == ref["n_moves"]).all(), n
        r1, r2 = bb(ref["final"])
        assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all(), n
        assert int(out["counters"][0]) == int(ref["n_moves"].sum()) and int(out["counters"][1]) == int(ref["n_turns"].sum()), n
        # the same games replayed from their move logs (rules-only instantiation) end on the same boards
        rep = engine.rollout_host(p1, p2, col, rng=Rng.replay_moves(out["moves"]))
        assert (rep["final_p1"] == r1).all() and (rep["final_p2"] == r2).all() and (rep["result"] == ref["results"]).all(), n


MANY_MOVES = [  # hill-climbed boards with 34 / 34 / 33 legal moves for colour 1 (more than the kernel's 32 scratch slots)
    "0120100002100220021201200021011001010120021220200212000000000021",
    "0000002112020000211122200220110000002120000001101201222011100000",
    "0000110020221202100110012021000000021220001011200210222101100000",
]


def test_arbitrary_boards_sample_vs_oracle(engine, cref, rollout_weights):
    """Single draws on arbitrary boards, including > 32 legal moves (the kernel's recompute path)."""
    from iago_b200 import Rng
    W, b = rollout_weights
    rng = np.random.default_rng(3)
    n = 3000
    st = np.zeros((n, 64), np.float32)
    many = np.array([[int(ch) for ch in s] for s in MANY_MOVES], np.float32)
    for i in range(n):
        if i % 3 == 0:
            st[i] = many[(i // 3) % 3]
            if i >= 9:  # perturb a few cells, keeps the move count high
                idx = rng.integers(0, 64, 2)
                st[i, idx] = rng.integers(0, 3, 2)
        else:
            fill = rng.random()
            r = rng.random(64)
            st[i] = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0))
    col = rng.integers(1, 3, n).astype(np.uint8)
    col[::3] = 1
    u = rng.random(n)
    u[:9] = [0.0, 0.999999999, 0.5, 1 - 2.0**-53, 0.25, 0.75, 0.1, 0.9, 0.33]
    p1, p2 = bb(st)
    got = engine.rollout_sample_host(p1, p2, col, rng=Rng.replay_uniforms(u))
    counts = []
    for i in range(n):
        counts.append(len(cref.legal_actions(st[i], int(col[i]))))
        assert int(got[i]) == cref.rollout_sample(st[i], int(col[i]), W, b, float(u[i]))
    assert max(counts) >= 34 and sum(c > 32 for c in counts) >= 3  # the recompute path was exercised


def test_arbitrary_boards_full_rollouts_vs_oracle(engine, cref, rollout_weights):
    """Whole rollouts from arbitrary (unreachable) boards in all three rng modes.  The paired kernel keeps 16 running sums per
    lane; the hill-climbed boards have more than 16 legal cells in one half of the board, so the first turn of those games takes
    its recompute path, and random dense boards exercise extreme u values at the seams of the two-sided cdf."""
    from iago_b200 import Rng
    W, b = rollout_weights
    rng = np.random.default_rng(11)
    n = 4000
    st = np.zeros((n, 64), np.float32)
    many = np.array([[int(ch) for ch in s] for s in MANY_MOVES], np.float32)
    for i in range(n):
        if i % 4 == 0:
            st[i] = many[(i // 4) % 3]
            if i >= 12:
                idx = rng.integers(0, 64, 2)
                st[i, idx] = rng.integers(0, 3, 2)
            if (i // 4) % 2:
                st[i] = st[i][::-1]   # the same board turned by 180 degrees: the crowded half changes lanes
        else:
            fill = rng.random()
            r = rng.random(64)
            st[i] = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0))
    col = rng.integers(1, 3, n).astype(np.uint8)
    col[::4] = 1
    half_counts = []
    for i in range(0, n, 4):
        acts = cref.legal_actions(st[i], int(col[i]))
        half_counts.append(max(sum(a < 32 for a in acts), sum(a >= 32 for a in acts)))
    assert sum(c > 16 for c in half_counts) >= 6
    p1, p2 = bb(st)
    ref = cref.simulate_batch(st, col.astype(np.int32), W, b, mode=cref.RNG_PHILOX, seed=77, game_id0=3, threads=0)
    out = engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=77, game_id0=3), want_moves=True)
    assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all()
    r1, r2 = bb(ref["final"])
    assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all()
    # replayed uniforms, with the extreme values in every position of the stream
    u = rng.random((n, 64))
    u[:, ::7] = np.array([0.0, 1 - 2.0**-53, 0.5, 2.0**-53, 0.999999, 1e-9, 0.25, 0.75, 0.125, 0.875])[rng.integers(0, 10, (n, 10))]
    ref = cref.simulate_batch(st, col.astype(np.int32), W, b, mode=cref.RNG_UNIFORMS, uniforms=u, threads=0)
    out = engine.rollout_host(p1, p2, col, rng=Rng.replay_uniforms(u), want_moves=True)
    assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all()
    # and the moves replayed
    out2 = engine.rollout_host(p1, p2, col, rng=Rng.replay_moves(ref["moves"]), want_moves=True)
    r1, r2 = bb(ref["final"])
    assert (out2["moves"] == ref["moves"]).all() and (out2["final_p1"] == r1).all() and (out2["final_p2"] == r2).all()


def test_arbitrary_placements_vs_oracle(engine, cref):
    """place_stone has no legality check (game.py:179-207): the replayed move is placed whatever it is — on an occupied cell, with
    nothing to flip, next to an edge.  20,000 random boards x random cells, both colours: the kernel's carry-propagation flips (four
    directions per lane, the second lane on the turned board) against the oracle's ray walks, board for board."""
    from iago_b200 import Rng
    rng = np.random.default_rng(23)
    n = 20000
    fill = rng.random((n, 1))
    r = rng.random((n, 64))
    st = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0)).astype(np.float32)
    col = rng.integers(1, 3, n).astype(np.uint8)
    forced = rng.integers(0, 64, (n, 64)).astype(np.int8)
    W, b = np.zeros((1, 2, 3, 3), np.float32), np.zeros(64, np.float32)   # unused in FORCED mode
    ref = cref.simulate_batch(st, col.astype(np.int32), W, b, mode=cref.RNG_FORCED, forced=forced, threads=0)
    p1, p2 = bb(st)
    out = engine.rollout_host(p1, p2, col, rng=Rng.replay_moves(forced), want_moves=True)
    r1, r2 = bb(ref["final"])
    assert (out["n_moves"] == ref["n_moves"]).all()
    assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all()
    assert (out["result"] == ref["results"]).all() and (out["moves"] == ref["moves"]).all()
    assert ref["n_moves"].sum() > 5 * n     # the games really placed stones


def test_logits_vs_oracle_and_reference(engine, cref, rollout_weights, golden_nets):
    W, b = rollout_weights
    g = golden_nets
    p1, p2 = bb(g["state"])
    got = engine.rollout_logits_host(p1, p2, g["color"].astype(np.uint8))
    for i in range(len(p1)):
        assert (got[i] == cref.rollout_logits(g["state"][i].astype(np.float32), int(g["color"][i]), W, b)).all()  # bit exact
    e = np.exp(got.astype(np.float64) - got.max(axis=1, keepdims=True))
    assert np.abs(e / e.sum(axis=1, keepdims=True) - g["rollout_prob"]).max() < 1e-6  # vs the reference's softmax output


def test_device_api_and_properties_at_full_size(engine):
    """Device-pointer entry point at 2^20 games: size-independent properties of a finished game."""
    import torch
    from iago_b200 import Rng, boards
    n = 1 << 20
    dev = torch.device("cuda", 0)
    p1 = torch.full((n,), boards.START_P1, dtype=torch.int64, device=dev)
    p2 = torch.full((n,), boards.START_P2, dtype=torch.int64, device=dev)
    col = torch.ones(n, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    out = engine.rollout(p1, p2, col, rng=Rng.philox(seed=1), counters=cnt)
    torch.cuda.synchronize()
    f1, f2 = out["final_p1"].cpu().numpy().view(np.uint64), out["final_p2"].cpu().numpy().view(np.uint64)
    assert (f1 & f2 == 0).all()                                     # no cell holds both colours
    pc = lambda x: np.array([bin(int(v)).count("1") for v in x[:20000]])
    n1, n2 = pc(f1), pc(f2)
    nm = out["n_moves"].cpu().numpy()
    assert (n1 + n2 == 4 + nm[:20000]).all()                         # one stone per placement
    assert (np.sign(n1 - n2) == out["result"].cpu().numpy()[:20000]).all()  # judge
    assert int(cnt[0]) == int(nm.sum())
    # terminal: neither side has a legal move on any final board
    lm1 = engine.legal_actions(out["final_p1"], out["final_p2"], col)
    lm2 = engine.legal_actions(out["final_p1"], out["final_p2"], col + 1)
    assert int((lm1 != 0).sum()) == 0 and int((lm2 != 0).sum()) == 0
    # determinism + sharding invariance: a slice computed alone equals the same game ids in the big batch
    sub = engine.rollout(p1[:4096], p2[:4096], col[:4096], rng=Rng.philox(seed=1, game_id0=65536))
    assert torch.equal(sub["final_p1"], out["final_p1"][65536:65536 + 4096])


def test_facade_simulate(engine, golden_simulate):
    from iago_b200.mcts_self_play import Simulate, simulate_batch
    from iago_b200 import boards
    g = golden_simulate
    for i in (0, 1, 1005):
        s = g["start"][i].reshape(8, 8).astype(np.float32)
        sim = Simulate(s, uniforms=g["uniforms"][i])
        assert sim.stone_num == int((s != 0).sum()) and sim.pass_flg is False
        r = sim(int(g["color"][i]))
        assert r == g["result"][i] and (sim.state.reshape(64).astype(np.uint8) == g["final"][i]).all()
        assert sim.moves == [int(a) for a in g["moves"][i] if a >= 0]
        assert sim.judge(int(g["color"][i])) == r
        assert (s.reshape(64).astype(np.uint8) == g["start"][i]).all()  # caller's board untouched (deepcopy)
    # step-wise API: turn() by turn() reproduces the same game
    i = 2
    sim = Simulate(g["start"][i].reshape(8, 8).astype(np.float32), uniforms=g["uniforms"][i])
    c = int(g["color"][i])
    while sim.stone_num < 64:
        sim.turn(c); sim.turn(3 - c)
    assert (sim.state.reshape(64).astype(np.uint8) == g["final"][i]).all()
    # default Philox path: deterministic per (seed, game id)
    out = simulate_batch(np.tile(boards.start_state(), (8, 1, 1)), 1)
    assert set(np.unique(out["result"])) <= {-1, 0, 1}


def test_empty_and_error_paths(engine):
    import torch
    import iago_b200
    from iago_b200 import Rng
    out = engine.rollout_host(np.zeros(0, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.uint8))
    assert out["result"].shape == (0,)
    full = np.array([0xFFFFFFFF00000000, 0], np.uint64), np.array([0x00000000FFFFFFFF, 0], np.uint64)
    out = engine.rollout_host(full[0], full[1], np.array([1, 2], np.uint8))   # full board / empty board
    assert out["n_moves"].tolist() == [0, 0] and out["result"].tolist() == [0, 0]
    with pytest.raises(iago_b200.IagoError):
        engine.rollout_host(full[0], full[1], np.ones(2, np.uint8), rng=Rng(mode=7))
    with pytest.raises(iago_b200.IagoError):
        engine.legal_actions(torch.zeros(4, dtype=torch.int32, device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda"),
                             torch.ones(4, dtype=torch.uint8, device="cuda"))


def test_fallback_sampler_and_host_chunking(engine, cref, rollout_weights):
    """(1) Weights whose tap sums could leave +-300 take the exp32 + fixed-point sampler in the library and in the oracle alike:
    trajectories stay bit exact.  (2) iago_rollout_host cuts large batches into pipelined chunks: sizes around the chunk
    boundaries (not multiples of 64, one game more / less than four CTAs' worth) give the oracle's games, and n = 0 is a no-op."""
    from iago_b200 import Rng, boards
    W, b = rollout_weights
    try:
        Wbig = (W * 40).astype(np.float32)
        assert not cref.policy_is_fast(Wbig, b)
        engine.load_rollout(Wbig, b)
        n = 4096
        p1, p2 = np.full(n, boards.START_P1, np.uint64), np.full(n, boards.START_P2, np.uint64)
        out = engine.rollout_host(p1, p2, np.ones(n, np.uint8), rng=Rng.philox(seed=5), want_moves=True)
        st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
        ref = cref.simulate_batch(st, 1, Wbig, b, mode=cref.RNG_PHILOX, seed=5, threads=0)
        assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all()
    finally:
        engine.load_rollout(W, b)
    for n in (0, 1, 63, 16383, 16384, 16385, 70001):
        p1, p2 = np.full(n, boards.START_P1, np.uint64), np.full(n, boards.START_P2, np.uint64)
        out = engine.rollout_host(p1, p2, np.ones(n, np.uint8), rng=Rng.philox(seed=9, game_id0=12345), want_moves=True)
        if n == 0:
            assert int(out["counters"][0]) == 0
            continue
        st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
        ref = cref.simulate_batch(st, 1, W, b, mode=cref.RNG_PHILOX, seed=9, game_id0=12345, threads=0)
        assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all(), n
        assert int(out["counters"][0]) == int(ref["n_moves"].sum()) and int(out["counters"][1]) == int(ref["n_turns"].sum()), n


def test_host_call_on_pinned_buffers(engine, cref, rollout_weights):
    """iago_rollout_host uses page-locked caller buffers in place: Philox games run as one launch that reads / writes the mapped
    buffers itself, replay streams go through direct async copies (two chunks from 16,384 games).  Both give the oracle's games,
    the same as the pageable (staged) path, with and without the optional outputs."""
    from iago_b200 import Rng, boards
    W, b = rollout_weights
    for n in (1, 63, 16385, 40000):
        p1, p2, col, out = engine.rollout_host_buffers(n, want_moves=True)
        p1[:], p2[:], col[:] = boards.START_P1, boards.START_P2, 1
        out["moves"][:] = 77
        engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=11, game_id0=5), out=out)
        st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
        ref = cref.simulate_batch(st, 1, W, b, mode=cref.RNG_PHILOX, seed=11, game_id0=5, threads=0)
        r1, r2 = cref.to_bitboards(ref["final"])
        assert (out["moves"] == ref["moves"]).all() and (out["result"] == ref["results"]).all(), n
        assert (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all() and (out["n_moves"] == ref["n_moves"]).all(), n
        assert int(out["counters"][0]) == int(ref["n_moves"].sum()) and int(out["counters"][1]) == int(ref["n_turns"].sum()), n
        # no optional outputs
        slim = dict(out, n_moves=None, moves=None, result=engine.pinned(n, np.int8), counters=np.zeros(2, np.uint64))
        engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=11, game_id0=5), out=slim)
        assert (slim["result"] == ref["results"]).all() and int(slim["counters"][0]) == int(ref["n_moves"].sum())
        # replayed moves from a pinned log: the direct-copy path
        forced = engine.pinned((n, 64), np.int8)
        forced[:] = ref["moves"]
        out2 = engine.rollout_host_buffers(n, want_moves=True)[3]
        engine.rollout_host(p1, p2, col, rng=Rng.replay_moves(forced), out=out2)
        assert (out2["moves"] == ref["moves"]).all() and (out2["final_p1"] == r1).all() and (out2["result"] == ref["results"]).all(), n
        # replayed uniforms from a pinned array (direct copies again), drawn so that the games differ from the Philox ones
        if n <= 16385:
            u = engine.pinned((n, 64), np.float64)
            u[:] = np.random.default_rng(n).random((n, 64))
            refu = cref.simulate_batch(st, 1, W, b, mode=cref.RNG_UNIFORMS, uniforms=u, threads=0)
            out4 = engine.rollout_host_buffers(n, want_moves=True)[3]
            engine.rollout_host(p1, p2, col, rng=Rng.replay_uniforms(u), out=out4)
            assert (out4["moves"] == refu["moves"]).all() and (out4["result"] == refu["results"]).all(), n
        # a pageable buffer among pinned ones falls back to staging
        out3 = dict(out2, result=np.empty(n, np.int8))
        engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=11, game_id0=5), out=out3)
        assert (out3["result"] == ref["results"]).all() and (out3["moves"] == ref["moves"]).all(), n


def test_streamed_host_batches_vs_oracle(engine, cref, rollout_weights):
    """iago_rollout_host_submit / _wait: several batches in flight on their own lanes (H2D copies, kernel, D2H copies per lane) give
    the oracle's games batch by batch — different sizes, seeds and game ids per lane, with and without the optional outputs, lanes
    reused — and the error paths (lane in flight, idle lane, pageable buffer, replay stream) report instead of running."""
    from iago_b200 import Rng, boards
    from iago_b200._lib import IagoError
    W, b = rollout_weights
    sizes = (4097, 1, 20000, 513)
    lanes = []
    for lane, n in enumerate(sizes):
        p1, p2, col, out = engine.rollout_host_buffers(n, want_moves=(lane % 2 == 0))
        p1[:], p2[:], col[:] = boards.START_P1, boards.START_P2, 1 + lane % 2
        if out["moves"] is not None:
            out["moves"][:] = 77
        lanes.append((p1, p2, col, out))
    for rnd in range(2):        # second round: the lanes and their device blocks are reused
        for lane, (p1, p2, col, out) in enumerate(lanes):
            engine.rollout_host_submit(lane, p1, p2, col, rng=Rng.philox(seed=21 + rnd, game_id0=1000 * lane), out=out)
        with pytest.raises(IagoError):
            engine.rollout_host_submit(0, *lanes[0][:3], rng=Rng.philox(seed=1), out=lanes[0][3])   # lane 0 is in flight
        for lane in reversed(range(len(sizes))):
            n = sizes[lane]
            out = engine.rollout_host_wait(lane)
            st = np.tile(boards.start_state().reshape(1, 64), (n, 1))
            ref = cref.simulate_batch(st, 1 + lane % 2, W, b, mode=cref.RNG_PHILOX, seed=21 + rnd, game_id0=1000 * lane, threads=0)
            r1, r2 = cref.to_bitboards(ref["final"])
            assert (out["result"] == ref["results"]).all() and (out["final_p1"] == r1).all() and (out["final_p2"] == r2).all(), (rnd, lane)
            assert (out["n_moves"] == ref["n_moves"]).all(), (rnd, lane)
            if out["moves"] is not None:
                assert (out["moves"] == ref["moves"]).all(), (rnd, lane)
            assert int(out["counters"][0]) == int(ref["n_moves"].sum()) and int(out["counters"][1]) == int(ref["n_turns"].sum()), (rnd, lane)
    with pytest.raises(IagoError):
        engine.rollout_host_wait(2)                                             # nothing in flight
    p1, p2, col, out = lanes[3]
    with pytest.raises(IagoError):
        engine.rollout_host_submit(1, p1, p2, col, rng=Rng.philox(seed=1), out=dict(out, result=np.empty(sizes[3], np.int8)))   # pageable
    with pytest.raises(IagoError):
        engine.rollout_host_submit(1, p1, p2, col, rng=Rng.replay_moves(np.zeros((sizes[3], 64), np.int8)), out=out)
    with pytest.raises(IagoError):
        engine.rollout_host_submit(engine.HOST_LANES, p1, p2, col, rng=Rng.philox(seed=1), out=out)
    # the lanes still work after the refused calls, and the synchronous call beside them gives the same games
    engine.rollout_host_submit(1, p1, p2, col, rng=Rng.philox(seed=5, game_id0=7), out=out)
    got = {k: (v.copy() if v is not None else None) for k, v in engine.rollout_host_wait(1).items()}
    sync = engine.rollout_host(p1, p2, col, rng=Rng.philox(seed=5, game_id0=7))
    assert (got["result"] == sync["result"]).all() and (got["final_p1"] == sync["final_p1"]).all() and (got["n_moves"] == sync["n_moves"]).all()


def test_facade_simulate_stream(engine):
    """mcts_self_play.simulate_stream: batches of different sizes streamed with three in flight are the games of one simulate_batch
    call over their concatenation (game ids run on across the batches)."""
    from iago_b200 import Rng, boards, mcts_self_play
    sizes = [700, 1, 4096, 33, 2500, 64, 900]
    total = sum(sizes)
    colors = (np.arange(total) % 2 + 1).astype(np.uint8)
    p1 = np.full(total, boards.START_P1, np.uint64)
    p2 = np.full(total, boards.START_P2, np.uint64)
    whole = mcts_self_play.simulate_batch(p1=p1, p2=p2, colors=colors, rng=Rng.philox(seed=9, game_id0=123), want_moves=True)
    cuts = np.cumsum([0] + sizes)
    parts = list(mcts_self_play.simulate_stream(((p1[a:b], p2[a:b], colors[a:b]) for a, b in zip(cuts[:-1], cuts[1:])),
                                                seed=9, game_id0=123, want_moves=True))
    assert [len(r["result"]) for r in parts] == sizes
    for k in ("result", "final_p1", "final_p2", "n_moves", "moves"):
        assert (np.concatenate([r[k] for r in parts]) == whole[k]).all(), k
    assert sum(int(r["counters"][0]) for r in parts) == int(whole["counters"][0])
    assert list(mcts_self_play.simulate_stream([], seed=1)) == []
