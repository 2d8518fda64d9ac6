"""CUDA rules kernels (through the C ABI) vs the reference goldens and the CPU oracle — bit exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bb(states):
    from iago_b200 import boards
    return boards.to_bitboards(states)


def test_legal_actions_golden(engine, golden_rules):
    g = golden_rules
    p1, p2 = bb(g["state"])
    got = engine.legal_actions_host(p1, p2, g["color"].astype(np.uint8))
    assert (got == g["legal_mask"]).all()


def test_place_stone_golden(engine, golden_rules):
    from iago_b200 import boards
    g = golden_rules
    p1, p2 = bb(g["ps_state"])
    q1, q2 = engine.place_stone_host(p1, p2, g["ps_action"], g["ps_color"].astype(np.uint8))
    assert (boards.from_bitboards(q1, q2).reshape(-1, 64).astype(np.uint8) == g["ps_after"]).all()
    # -1 = pass = no-op
    q1, q2 = engine.place_stone_host(p1[:10], p2[:10], np.full(10, -1, np.int8), np.ones(10, np.uint8))
    assert (q1 == p1[:10]).all() and (q2 == p2[:10]).all()


def test_random_boards_vs_oracle(engine, cref):
    """Arbitrary (mostly unreachable) boards at every density, both colours, legal + every-cell placement."""
    from iago_b200 import boards
    rng = np.random.default_rng(7)
    n = 4000
    fill = rng.random((n, 1))
    r = rng.random((n, 64))
    st = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0)).astype(np.float32)
    col = rng.integers(1, 3, n).astype(np.uint8)
    p1, p2 = bb(st)
    got = engine.legal_actions_host(p1, p2, col)
    for i in range(n):
        m = 0
        for a in cref.legal_actions(st[i], int(col[i])):
            m |= 1 << a
        assert int(got[i]) == m
    act = rng.integers(0, 64, n).astype(np.int8)  # no legality check: any cell, occupied or not
    q1, q2 = engine.place_stone_host(p1, p2, act, col)
    after = boards.from_bitboards(q1, q2)
    for i in range(n):
        t = st[i].reshape(8, 8).copy()
        cref.place_stone(t, int(act[i]), int(col[i]))
        assert (after[i] == t).all()


def test_perft_on_gpu(engine):
    """Breadth-first perft with the batched kernels: 4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288."""
    import torch
    from iago_b200 import boards
    dev = torch.device("cuda", 0)
    p1 = torch.tensor([boards.START_P1], dtype=torch.int64, device=dev)
    p2 = torch.tensor([boards.START_P2], dtype=torch.int64, device=dev)
    col = torch.ones(1, dtype=torch.uint8, device=dev)
    passed = torch.zeros(1, dtype=torch.bool, device=dev)
    done = 0  # leaves that ended early (double pass) count once per remaining depth, like the recursive perft
    counts = []
    for depth in range(9):
        legal = engine.legal_actions(p1, p2, col)
        bits = ((legal.unsqueeze(1) >> torch.arange(64, device=dev)) & 1).bool()
        nmv = bits.sum(1)
        has = nmv > 0
        # boards with moves expand; boards without: pass once (if not already passed) or terminate
        idx, act = bits.nonzero(as_tuple=True)
        c1, c2, cc = p1[idx].clone(), p2[idx].clone(), col[idx].clone()
        engine.place_stone(c1, c2, act.to(torch.int8), cc)
        pas = (~has) & (~passed)
        term = (~has) & passed
        done += int(term.sum())
        p1 = torch.cat([c1, p1[pas]]); p2 = torch.cat([c2, p2[pas]])
        col = torch.cat([3 - cc, 3 - col[pas]])
        passed = torch.cat([torch.zeros(len(c1), dtype=torch.bool, device=dev),
                            torch.ones(int(pas.sum()), dtype=torch.bool, device=dev)])
        counts.append(len(p1) + done)
    assert counts == [4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288]


def test_facade_gamefunctions(engine):
    from iago_b200 import boards
    from iago_b200.game import GameFunctions as gf
    s = boards.start_state()
    assert gf.legal_actions(s, 1) == [19, 26, 37, 44]
    assert gf.legal_actions(s, 2) == [20, 29, 34, 43]
    r = gf.place_stone(s, 19, 1)
    assert r is s and s[2, 3] == 1 and s[3, 3] == 1 and s[4, 4] == 2
    assert gf.place_stone(s, -1, 2) is s
    x = gf.make_state_var(s, 2)
    assert x.shape == (1, 2, 8, 8) and x.dtype == np.float32 and x[0, 1, 4, 4] == 1 and x[0, 0, 3, 3] == 1
    assert gf.ac2pos([0, 63]) == [[1, 1], [8, 8]] and gf.is_outside([8, 0]) and not gf.is_outside([7, 7])
