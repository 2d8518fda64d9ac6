"""game.Game (auto play: SLPolicy as colour 1 vs PV-MCTS as colour 2, game.py:75-142) and the masked sampler behind
get_action_auto (iago_sample_masked) against numpy's np.random.choice on the same probabilities."""
import numpy as np
import pytest

from conftest import model_file

pytestmark = pytest.mark.gpu


def test_masked_sampler_equals_np_random_choice(engine, golden_nets):
    import torch
    from iago_b200 import Rng, boards, network
    sl = network.SLPolicy().load(model_file("sl_model.npz"))
    g = golden_nets
    n = 256
    st, col = g["state"][:n].reshape(n, 8, 8), g["color"][:n]
    p1, p2 = boards.to_bitboards(st)
    own = np.where(col == 1, p1, p2)
    opp = np.where(col == 1, p2, p1)
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a.view(np.int64).copy()).to(dev)
    probs = engine.policy_forward(sl.slot, t(p1), t(p2), torch.from_numpy(col.astype(np.uint8)).to(dev), probs=True)
    u = np.random.RandomState(5).random_sample(n)
    draws = torch.zeros(n, dtype=torch.int32, device=dev)
    act = engine.sample_masked(probs, t(own), t(opp), draws, Rng.replay_uniforms(torch.from_numpy(u.reshape(n, 1)).to(dev))).cpu().numpy()
    pr = probs.cpu().numpy()
    same = 0
    for i in range(n):
        legal = boards.mask_to_actions(g["legal_mask"][i])
        if not legal:
            assert act[i] == -1
            same += 1
            continue
        valid = np.zeros(64); valid[legal] = 1
        p = pr[i] * valid
        cdf = np.cumsum(p / np.sum(p)); cdf /= cdf[-1]
        want = int(np.searchsorted(cdf, u[i], side="right"))           # == np.random.choice(64, p=p/sum(p)) for this uniform
        assert act[i] in legal
        same += int(act[i] == want)
    assert same >= n - 1                                                 # a draw within ~1e-16 of a cdf edge may round the other way
    assert int(draws.sum()) == int((g["legal_mask"][:n] != 0).sum())     # exactly one uniform per move, none when there is no move


def test_auto_game_runs_to_the_end(engine, tmp_path, monkeypatch):
    from iago_b200.game import Game, GameFunctions
    monkeypatch.chdir(tmp_path)
    game = Game(True, verbose=False, n_playouts=192, leaf_batch=16, seed=4)
    turns = 0
    while game.stone_num < 64 and turns < 140:
        game.turn(1, True)
        game.turn(2, True)
        turns += 2
    assert game.stone_num >= 64
    assert GameFunctions.legal_actions(game.state, 1) == [] or game.stone_num == 64
    jd = game.judge()
    assert jd.startswith("IaGo(SLPolicy):") and "IaGo(PV-MCTS):" in jd
    game.gamelog += jd + "\n"
    game.save_gamelog()
    files = list((tmp_path / "gamelog").iterdir())
    assert len(files) == 1 and files[0].read_text().startswith("IaGo \n")
    assert game.gamelog.count("]IaGo(PV-MCTS): ") >= 20
