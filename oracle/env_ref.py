"""numpy restatement of rl_env.GameEnv (rl_env.py:10-185) — TEST INFRASTRUCTURE ONLY.

step (rl_env.py:41-74): the learner (colour 1) plays `action` (an illegal one is replaced; the reference uses Python's
random.choice there, this restatement — like the product — takes positions[floor(u * len)] from the injected stream),
then the opponent (colour 2) answers through get_position (rl_env.py:152-172): p = out - min(out) over all 64 cells,
np.random.choice, re-drawn until legal.  Pinned against the UNMODIFIED reference class run under the chainer stand-in
(tests/golden/env.npz, oracle/gen_golden.py gen_env).  Rules come from the C oracle; the opponent's probabilities are
injected (policy_func(state) -> float32[64] for colour 2 to move).
"""
import numpy as np

from . import cref


def choice_unmasked(prob, u):
    """np.random.choice(64, p=(prob - min)/sum) for the uniform u (rl_env.py:166-167)."""
    p = np.array(prob, np.float32).copy()
    p -= np.min(p)
    p = p / np.sum(p)
    cdf = np.cumsum(p.astype(np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


class RefEnv:
    def __init__(self, policy_func, uniforms):
        self.policy_func, self.uniforms = policy_func, np.asarray(uniforms, np.float64)
        self.reset()

    def reset(self):
        self.state = cref.start_board()
        self.stone_num, self.pass_flg, self.draws = 4, False, 0
        self.opp_actions = []

    def _u(self):
        u = self.uniforms[self.draws]
        self.draws += 1
        return u

    def step(self, action):
        done = False
        acts = cref.legal_actions(self.state, 1)
        if acts:
            if action not in acts:
                action = acts[min(int(self._u() * len(acts)), len(acts) - 1)]
            cref.place_stone(self.state, action, 1)
            self.stone_num += 1
            self.pass_flg = False
        else:
            if self.pass_flg:
                done = True
            self.pass_flg = True
        acts = cref.legal_actions(self.state, 2)
        if acts:
            prob = self.policy_func(self.state)
            while True:
                idx = choice_unmasked(prob, self._u())
                if idx in acts:
                    break
            cref.place_stone(self.state, idx, 2)
            self.opp_actions.append(idx)
            self.stone_num += 1
            self.pass_flg = False
        else:
            self.opp_actions.append(-1)
            if self.pass_flg:
                done = True
            self.pass_flg = True
        if self.stone_num >= 64:
            done = True
        return done

    def judge(self):
        a, b = int((self.state == 1).sum()), int((self.state == 2).sum())
        return (a > b) - (a < b)
