"""Module-level drop-in: the UNMODIFIED reference MCTS.py (the copy oracle/fetch_ref.py puts under baseline/_ref) is imported with
this package's modules standing in for the ones it imports — `network`, `mcts_self_play`, `game` — and a four-line `chainer` adapter
(serializers.load_npz -> model.load), i.e. exactly the binding INTEGRATION.md describes.  Its own playout loop, Node arithmetic and
np.random seeds then drive the GPU nets and rollouts, and must rebuild the trees it built on its own stack (tests/golden/mcts.npz):
the rollout result z of every playout identical, the value v within the net tolerance, visit counts identical, Q within 2e-4."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from conftest import MODELS, load_golden

pytestmark = pytest.mark.gpu


def flatten(root):
    nodes, parent, action, i = [root], [-1], [0], 0
    while i < len(nodes):
        for a, ch in nodes[i].children.items():
            nodes.append(ch); parent.append(i); action.append(int(a))
        i += 1
    return (np.array(parent), np.array(action), np.array([nd.n_visits for nd in nodes]), np.array([float(nd.Q) for nd in nodes]),
            np.array([float(nd.P) for nd in nodes]))


@pytest.fixture(scope="module")
def ref_mcts(engine):
    ref_dir = os.path.dirname(MODELS)
    path = os.path.join(ref_dir, "MCTS.py")
    if not os.path.isfile(path):
        pytest.fail(f"{path} missing: run `python oracle/fetch_ref.py` in the build container")
    from iago_b200 import game, mcts_self_play, network
    chainer = types.ModuleType("chainer")
    chainer.config = types.SimpleNamespace(train=False, enable_backprop=False)
    chainer.serializers = types.SimpleNamespace(load_npz=lambda p, model, *a, **k: model.load(p))
    chainer.cuda = chainer.optimizers = types.SimpleNamespace()
    chainer.Variable = network.Variable
    saved = {k: sys.modules.get(k) for k in ("chainer", "network", "mcts_self_play", "game")}
    sys.modules.update(chainer=chainer, network=network, mcts_self_play=mcts_self_play, game=game)
    cwd = os.getcwd()
    os.chdir(ref_dir)                                   # the reference opens './models/*.npz'
    mcts_self_play.USE_NUMPY_RNG = True
    try:
        spec = importlib.util.spec_from_file_location("MCTS_reference_file", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.Node.copy = lambda self: self               # MCTS.py:106 calls a method Node does not have (oracle/ref_harness.py)
        yield mod
    finally:
        mcts_self_play.USE_NUMPY_RNG = False
        os.chdir(cwd)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("case", ["after19", "mid30", "late52_lam1_thr2", "late56_lam0_thr1"])
def test_reference_mcts_file_rebuilds_its_trees_on_the_gpu_modules(ref_mcts, case):
    g = load_golden("mcts")
    f = lambda k: g[f"{case}/{k}"]
    m = ref_mcts.MCTS(lmbda=float(f("lmbda")), c_puct=float(f("c_puct")), n_thr=int(f("n_thr")))
    zs, vs = [], []
    vf, rf = m.value_func, m.evaluate_rollout
    m.value_func = lambda st, c: vs.append(vf(st, c)) or vs[-1]
    m.evaluate_rollout = lambda st, c: zs.append(rf(st, c)) or zs[-1]
    state = f("root_state").reshape(8, 8).astype(np.float32)
    for k in range(int(f("n_playouts"))):
        np.random.seed(7000 + k)                        # the generator's seeds (oracle/gen_golden.py gen_mcts)
        n_z = len(zs)
        m.playout(state.copy(), int(f("root_color")), m.root)
        if float(f("lmbda")) > 0:
            assert len(zs) == n_z + 1 and zs[-1] == int(f("z")[k]), (case, k)
    if float(f("lmbda")) < 1:
        assert np.abs(np.array(vs, np.float64) - f("v")[:len(vs)]).max() <= 2e-4
    parent, action, n, Q, P = flatten(m.root)
    assert (parent == f("tree_parent")).all() and (action == f("tree_action")).all()
    assert (n == f("tree_n")).all()
    assert np.abs(Q - f("tree_Q")).max() <= 2e-4 and np.abs(P - f("tree_P")).max() <= 2e-4
    best = max(m.root.children.items(), key=lambda an: an[1].n_visits)[0]
    assert best == int(f("best"))
