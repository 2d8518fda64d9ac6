"""oracle/env_ref.py pinned against the UNMODIFIED rl_env.GameEnv (tests/golden/env.npz): same opponent answers, same number
of np.random uniforms consumed, same final boards and judge when fed the reference's own probabilities and uniforms."""
import numpy as np

from conftest import load_golden


def test_env_restatement_reproduces_reference_games(cref):
    from oracle import env_ref
    g = load_golden("env")
    for i in range(len(g["seed"])):
        probs = iter(g["probs"][i])
        env = env_ref.RefEnv(lambda st: next(probs), g["uniforms"][i])
        done = False
        for k in range(int(g["n_steps"][i])):
            assert not done
            before = env.state.copy()
            done = env.step(int(g["actions"][i][k]))
            if env.opp_actions[-1] < 0:
                next(probs, None)   # the generator logged a forward for this step although the opponent had to pass
        assert done
        assert env.opp_actions == g["opp_actions"][i][:len(env.opp_actions)].tolist()
        assert env.draws == int(g["n_draws"][i])
        assert (env.state.reshape(64).astype(np.uint8) == g["final"][i]).all()
        assert env.judge() == int(g["judge"][i])


def test_choice_unmasked_is_numpy_choice():
    from oracle import env_ref
    rng = np.random.default_rng(0)
    for t in range(300):
        p = rng.random(64).astype(np.float32) ** 4
        q = p - p.min()
        a = np.random.RandomState(t).choice(64, p=q / np.sum(q))
        assert a == env_ref.choice_unmasked(p, np.random.RandomState(t).random_sample())
