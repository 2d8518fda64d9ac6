"""Value-data generation (iago_value_selfplay through the C ABI, iago_b200/value_self_play.py) vs the UNMODIFIED
value_self_play.SelfPlay games in tests/golden/valuegen.npz and vs the numpy restatement oracle/valuegen_ref.py."""
import numpy as np
import pytest

from conftest import load_golden, model_file

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(engine):
    from iago_b200 import network
    return network.SLPolicy().load(model_file("sl_model.npz")), network.SLPolicy().load(model_file("rl_model.npz"))


def test_reference_games_replayed(engine, models):
    """The reference's own games under the uniforms its np.random seed produced: recorded position, result, final board and the
    number of uniforms consumed are identical."""
    from iago_b200 import Rng
    from iago_b200.value_self_play import play_games
    g = load_golden("valuegen")
    out = play_games(g["stop_num"], *models, rng=Rng.replay_uniforms(g["uniforms"]))
    assert (out["state"].reshape(-1, 64).astype(np.uint8) == g["state"]).all()
    assert (out["result"] == g["result"]).all()
    assert (out["final"].reshape(-1, 64).astype(np.uint8) == g["final"]).all()
    assert (out["draws"] == g["n_draws"]).all()


def test_matches_oracle_on_random_stop_nums(engine, models):
    """Random stop_num in [4, 64) and random uniforms, live nets on both sides (GPU trunk vs fp32 numpy forward).  The two
    forwards differ by ~1e-4 in probability, so a draw within that distance of a cdf edge may pick the neighbouring cell and
    the games then part ways; at least 90 % of the games must be identical in every output."""
    from iago_b200 import Rng
    from iago_b200.value_self_play import play_games
    from oracle import nets, valuegen_ref
    rs = np.random.RandomState(5)
    n = 40
    stop = rs.randint(4, 64, size=n)
    u = rs.random_sample((n, 200))
    out = play_games(stop, *models, rng=Rng.replay_uniforms(u))
    psl, prl = nets.load_params(model_file("sl_model.npz")), nets.load_params(model_file("rl_model.npz"))
    f = lambda p: (lambda st, c: nets.sl_logits(p, nets.planes_from_state(st[None], c))[0])
    same = 0
    for i in range(n):
        r = valuegen_ref.play(int(stop[i]), f(psl), f(prl), u[i])
        ok = ((out["state"][i] == r["state"]).all() and out["result"][i] == r["result"] and out["color"][i] == r["color"]
              and out["action"][i] == r["action"] and (out["final"][i] == r["final"]).all() and out["draws"][i] == r["draws"])
        same += bool(ok)
    print(f"value-data games identical to the oracle: {same} of {n}")
    assert same >= 0.9 * n


def test_invariants_philox(engine, models):
    from iago_b200 import Rng, boards
    from iago_b200.engine import STREAM_VALUEGEN
    from iago_b200.value_self_play import play_games
    from oracle import cref
    rs = np.random.RandomState(9)
    n = 2048
    stop = rs.randint(4, 64, size=n)
    rng = Rng.philox(seed=3, game_id0=100, stream_id=STREAM_VALUEGEN)
    a, b = play_games(stop, *models, rng=rng), play_games(stop, *models, rng=rng)
    for k in ("state", "result", "color", "action", "final", "draws"):
        assert (a[k] == b[k]).all(), k                                  # deterministic
    stones = (a["state"] != 0).sum(axis=(1, 2))
    played = a["action"] >= 0
    assert (stones[played] == stop[played]).all()                       # recorded exactly when stone_num reached stop_num
    assert (a["result"][~played] == -1).all()
    assert set(np.unique(a["result"])) <= {-1, 0, 1} and set(np.unique(a["color"])) <= {1, 2}
    for i in np.nonzero(played)[0][:200]:
        # the recorded board is the mover's view (mover 2, other side 1): the random move must be legal for "2" there
        assert int(a["action"][i]) in cref.legal_actions(a["state"][i], 2)
    # a sub-batch with shifted game ids plays the same games (results do not depend on batch composition)
    c = play_games(stop[512:1024], *models, rng=Rng.philox(seed=3, game_id0=100 + 512, stream_id=STREAM_VALUEGEN))
    assert (c["final"] == a["final"][512:1024]).all() and (c["result"] == a["result"][512:1024]).all()
