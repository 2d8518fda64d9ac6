"""oracle/mcts_ref.py pinned against the reference's own search trees (tests/golden/mcts.npz, produced by the UNMODIFIED
MCTS.playout through oracle/gen_golden.py) by replaying the logged leaf values, rollout results and priors."""
import numpy as np
import pytest

from conftest import load_golden


def replay_search(g, name, cref):
    from oracle import mcts_ref
    state = g[f"{name}/root_state"].astype(np.float32).reshape(8, 8)
    color = int(g[f"{name}/root_color"])
    v, z = g[f"{name}/v"], g[f"{name}/z"]
    priors = {(s.tobytes(), int(c)): p for s, c, p in zip(g[f"{name}/prior_state"], g[f"{name}/prior_color"], g[f"{name}/prior"])}
    box = {}

    def value_func(st, c):
        return np.float32(v[box["s"].done])

    def rollout_func(st, c, k):
        return int(z[k])

    def policy_func(st, c):
        return priors[(st.astype(np.uint8).reshape(64).tobytes(), int(c))]

    lm, cp = float(g[f"{name}/lmbda"]), float(g[f"{name}/c_puct"])
    # the reference's defaults are the Python int 1 / float 0.5; keep an int c_puct an int so the scalar types match
    s = mcts_ref.RefSearch(state, color, value_func, rollout_func, policy_func, lmbda=lm,
                           c_puct=int(cp) if cp == int(cp) else cp, n_thr=int(g[f"{name}/n_thr"]))
    box["s"] = s
    return s


CASES = ["after19", "mid30", "late52_lam1_thr2", "late56_lam0_thr1"]


@pytest.mark.parametrize("name", CASES)
def test_sequential_restatement_reproduces_reference_tree(cref, name):
    g = load_golden("mcts")
    s = replay_search(g, name, cref)
    for _ in range(int(g[f"{name}/n_playouts"])):
        s.playout_sequential()
    t = s.flatten()
    assert (t["parent"] == g[f"{name}/tree_parent"]).all()
    assert (t["action"] == g[f"{name}/tree_action"]).all()
    assert (t["n"] == g[f"{name}/tree_n"]).all()
    assert (t["Q"] == g[f"{name}/tree_Q"]).all()      # bit-exact: same scalar types as the reference
    assert (t["P"] == g[f"{name}/tree_P"]).all()
    assert s.best_move() == int(g[f"{name}/best"])


@pytest.mark.parametrize("name", CASES)
def test_batched_restatement_with_batch_one_is_the_reference(cref, name):
    g = load_golden("mcts")
    s = replay_search(g, name, cref)
    s.search(int(g[f"{name}/n_playouts"]), leaf_batch=1)
    t = s.flatten()
    for k in ("parent", "action", "n", "Q", "P"):
        assert (t[k] == g[f"{name}/tree_{k}"]).all(), k


def test_batched_restatement_is_consistent(cref):
    """leaf_batch > 1: every playout is backed up exactly once, virtual visits vanish, root visits = playouts."""
    g = load_golden("mcts")
    s = replay_search(g, "after19", cref)
    s.cache_value = False
    s.value_func = lambda st, c: np.float32(0.1 * ((int(st.sum()) % 7) - 3))      # any deterministic evaluator will do here
    s.rollout_func = lambda st, c, k: (k * 7919 % 3) - 1
    s.policy_func = lambda st, c: (np.arange(64, dtype=np.float32) % 5 + 1) / np.float32(192)
    s.search(384, leaf_batch=16, virtual_loss=1.0)
    t = s.flatten(exact=False)
    assert t["n"][0] == 384
    stack = [s.root]
    while stack:
        nd = stack.pop()
        assert nd.vn == 0 and nd.pending is None
        if nd.children:
            assert sum(ch.n_visits for ch in nd.children.values()) <= nd.n_visits
        stack.extend(nd.children.values())
