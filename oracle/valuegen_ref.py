"""numpy restatement of value_self_play.SelfPlay (value_self_play.py:12-162) — TEST INFRASTRUCTURE ONLY.

One game: the SL policy plays both colours while stone_num < stop_num (:35-37); the side to move records the board from
its own view (:39-44, mover's stones 2, the other side's 1) and plays one uniformly random legal move (:46-53; no legal
move -> (state, -1)); the RL policy plays the game out (:55-57); result = judge(color) (:59).  get_position (:131-149) is
np.random.choice(64, p=softmax(net output)) with an illegal cell replaced by random.choice(positions).

Randomness is injected: uniform k of the game's stream feeds np.random.choice (one per sampled move), and both
random.choice calls take positions[floor(u * len)] of the NEXT uniform (the reference uses Python's own generator there;
oracle/gen_golden.py routes it to the same stream when it runs the unmodified class).  The reference's softmax does not
subtract the maximum; this one does (same distribution wherever the reference's exp is finite — it raises above 88.7).
Pinned by tests/golden/valuegen.npz: games of the UNMODIFIED value_self_play.SelfPlay run under the chainer stand-in with
a stand-in for the deleted SLPolicy module (SLPolicyNet = network.SLPolicy without its final softmax).
Rules come from the C oracle (oracle/othello_ref.c); `logits_sl(state, color)` / `logits_rl` return float32[64].
"""
import numpy as np

from . import cref


def softmax_choice(logits, u):
    x = np.asarray(logits, np.float32)
    ex = np.exp(x - x.max())
    p = ex / np.sum(ex)
    cdf = np.cumsum(p.astype(np.float64))
    cdf /= cdf[-1]
    return int(np.searchsorted(cdf, u, side="right"))


def pick(seq, u):
    return seq[min(int(u * len(seq)), len(seq) - 1)]


def play(stop_num, logits_sl, logits_rl, uniforms):
    state = cref.start_board()
    stone_num, pass_flg, draws = 4, False, 0
    uniforms = np.asarray(uniforms, np.float64)

    def turn(cl, logits_fn):
        nonlocal stone_num, pass_flg, draws
        acts = cref.legal_actions(state, cl)
        if acts:
            a = softmax_choice(logits_fn(state, cl), uniforms[draws]); draws += 1
            if a not in acts:
                a = pick(acts, uniforms[draws]); draws += 1
            cref.place_stone(state, a, cl)
            pass_flg = False
            stone_num += 1
        else:
            if pass_flg:
                stone_num = 64
            pass_flg = True

    cl = 1
    while stone_num < stop_num:
        turn(cl, logits_sl)
        cl = 3 - cl
    color = cl
    rec = state.copy()
    if color == 1:
        rec = rec * (3 - rec) * (3 - rec) / 2
    acts = cref.legal_actions(state, cl)
    if not acts:
        return dict(state=rec.astype(np.float32), result=-1, color=color, action=-1, final=state, draws=draws)
    a = pick(acts, uniforms[draws]); draws += 1
    cref.place_stone(state, a, cl)
    pass_flg = False
    stone_num += 1
    cl = 3 - cl
    while stone_num < 64:
        turn(cl, logits_rl)
        cl = 3 - cl
    me, op = int((state == color).sum()), int((state == 3 - color).sum())
    return dict(state=rec.astype(np.float32), result=(me > op) - (me < op), color=color, action=a, final=state, draws=draws)
