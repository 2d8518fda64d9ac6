"""iago_b200/load.py vs the UNMODIFIED reference load.py (tests/golden/load.npz: the npy files its main() wrote for a synthetic
data.txt under a fixed np.random seed, and its rotate / transpose tables)."""
import os

import numpy as np

from conftest import load_golden


def test_action_maps():
    from iago_b200 import load
    g = load_golden("load")
    a = np.arange(64.0)
    assert (load.rotate(a) == g["rotate"]).all() and (load.transpose(a) == g["transpose"]).all()
    assert sorted(load.rotate(a)) == list(a)                      # permutations of the board


def test_main_writes_the_reference_files(tmp_path):
    from iago_b200 import load
    g = load_golden("load")
    (tmp_path / "txt").mkdir(); (tmp_path / "npy").mkdir()
    with open(tmp_path / "txt" / "data.txt", "w") as f:
        f.writelines([str(l) for l in g["lines"]])
    np.random.seed(int(g["seed"]))
    load.main(txt=str(tmp_path / "txt" / "data.txt"), out_dir=str(tmp_path / "npy"))
    for name in ("states", "actions", "states_test", "actions_test"):
        got = np.load(tmp_path / "npy" / f"{name}.npy")
        assert got.shape == g[name].shape and (got == g[name]).all(), name


def test_augmentation_moves_the_action_with_the_board():
    """Mark the action cell on the board: after each of the 8 transforms the mark must sit at the transformed action."""
    from iago_b200 import load
    rs = np.random.RandomState(0)
    acts = rs.randint(0, 64, size=16).astype(np.float64)
    st = np.zeros((16, 8, 8))
    st[np.arange(16), acts.astype(int) // 8, acts.astype(int) % 8] = 7
    S, A = load.augment(st, acts)
    assert S.shape == (128, 8, 8)
    for s, a in zip(S, A):
        assert s.reshape(64)[int(a)] == 7 and s.sum() == 7
