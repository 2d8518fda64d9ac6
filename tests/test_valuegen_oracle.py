"""oracle/valuegen_ref.py pinned against the UNMODIFIED value_self_play.SelfPlay (tests/golden/valuegen.npz, produced by
oracle/gen_golden.py gen_valuegen): fed the reference's own net outputs and uniforms, the restatement records the same
position, plays the same game to the same final board and result, and consumes the same number of uniforms."""
import numpy as np

from conftest import load_golden


def test_valuegen_restatement_reproduces_reference_games(cref):
    from oracle import valuegen_ref
    g = load_golden("valuegen")
    assert len(g["seed"]) >= 20
    for i in range(len(g["seed"])):
        it = iter(g["logits"][i][:int(g["n_logits"][i])])
        fn = lambda state, color: next(it)
        r = valuegen_ref.play(int(g["stop_num"][i]), fn, fn, g["uniforms"][i])
        assert next(it, None) is None                      # as many forwards as the reference ran
        assert (r["state"].reshape(64).astype(np.uint8) == g["state"][i]).all()
        assert r["result"] == int(g["result"][i])
        assert r["draws"] == int(g["n_draws"][i])
        assert (r["final"].reshape(64).astype(np.uint8) == g["final"][i]).all()


def test_softmax_choice_is_numpy_choice():
    from oracle import valuegen_ref
    rng = np.random.default_rng(1)
    for t in range(300):
        x = (rng.standard_normal(64) * 5).astype(np.float32)
        ex = np.exp(x)
        a = np.random.RandomState(t).choice(64, p=ex / np.sum(ex))      # value_self_play.py:143, :171-173
        assert a == valuegen_ref.softmax_choice(x, np.random.RandomState(t).random_sample())
