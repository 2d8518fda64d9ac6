"""numpy restatement of src/rl_self_play.py:8-145 — TEST INFRASTRUCTURE ONLY.

Plays n games in lockstep exactly as `Game(model1, model2)()` would play each of them: turn(1) with model1 then
turn(2) with model2 while stone_num < 64 (rl_self_play.py:27-31), get_action = float32 policy probabilities times a
float64 validity mask, renormalised, inverse cdf against one uniform (np.random.choice, rl_self_play.py:111-127),
records of the learner's swapped pre-move board and action (:134-138), judge from colour 1's view (:91-100).
Rules come from the C oracle (oracle/othello_ref.c), the nets from oracle/nets.py.
"""
import numpy as np

from . import cref, nets


def play(p_learner, p_opponent, init_states, uniforms=None, greedy=False, dtype=np.float32):
    states = np.array(init_states, np.float32).reshape(-1, 8, 8).copy()
    n = len(states)
    stone_num = np.full(n, 4)          # rl_self_play.py:20
    pass_flg = np.zeros(n, bool)
    placed = np.zeros(n, int)
    rec_states = [[] for _ in range(n)]
    rec_actions = [[] for _ in range(n)]
    moves = [[] for _ in range(n)]
    while (stone_num < 64).any():
        for color, params in ((1, p_learner), (2, p_opponent)):
            act = [g for g in range(n) if stone_num[g] < 64]
            if not act:
                break
            x = nets.planes_from_state(states[act], color, dtype)
            logits = nets.sl_logits(params, x)
            prob = nets.softmax(logits.copy()).astype(np.float32)
            for j, g in enumerate(act):
                actions = cref.legal_actions(states[g], color)
                if actions:
                    if greedy:
                        a = actions[int(np.argmax(logits[j][actions]))]   # lowest index on ties
                    else:
                        valid = np.zeros(64)
                        valid[actions] = 1
                        p = prob[j] * valid
                        p = p / np.sum(p)
                        cdf = np.cumsum(p)
                        cdf /= cdf[-1]
                        a = int(np.searchsorted(cdf, uniforms[g][placed[g]], side="right"))
                    if color == 1:
                        s = states[g]
                        rec_states[g].append((s * (3 - s) * (3 - s) / 2).astype(np.float32))
                        rec_actions[g].append(a)
                    cref.place_stone(states[g], a, color)
                    moves[g].append(a)
                    placed[g] += 1
                    pass_flg[g] = False
                    stone_num[g] += 1
                else:
                    if pass_flg[g]:
                        stone_num[g] = 64
                    pass_flg[g] = True
    n1 = (states == 1).sum(axis=(1, 2))
    n2 = (states == 2).sum(axis=(1, 2))
    return dict(final=states, result=np.sign(n1 - n2).astype(np.int8), rec_states=rec_states, rec_actions=rec_actions, moves=moves)
