"""Game — drop-in for /root/reference/src/rl_self_play.py:8-145 on the batched self-play engine.

    game = Game(model1, model2)          # network.SLPolicy objects; model1 = learner (colour 1), model2 = opponent
    game.state[r, c] = 2                 # callers may poke the start board (src/train_rl.py:43-46)
    states, actions, judge = game()      # learner's swapped pre-move boards (list of (8,8) float32), actions, +1/0/-1

`play_games` is the batched form (n games per launch sequence) that train_rl-style loops should use.
"""
import itertools

import numpy as np
import torch

from . import boards
from .engine import RNG_UNIFORMS, Rng, STREAM_SELFPLAY, default_engine

_game_counter = itertools.count()


def play_games(model1, model2, n, init_states=None, *, greedy=False, rng=None, rec_cap=40, want_moves=False, device=0):
    """n lockstep games. init_states: optional (n,8,8) start boards. Returns numpy dict:
    final (n,8,8), result int8[n], rec_states list per game of swapped boards, rec_own/rec_opp/rec_action/n_rec, moves."""
    eng = default_engine(device)
    dev = torch.device("cuda", device)
    i1 = i2 = None
    if init_states is not None:
        q1, q2 = boards.to_bitboards(init_states)
        i1 = torch.from_numpy(q1.view(np.int64).copy()).to(dev)
        i2 = torch.from_numpy(q2.view(np.int64).copy()).to(dev)
    if rng is not None and rng.mode == RNG_UNIFORMS and not torch.is_tensor(rng.uniforms):
        rng = Rng.replay_uniforms(torch.from_numpy(np.ascontiguousarray(rng.uniforms, np.float64).reshape(n, -1)).to(dev))
    if rng is not None and rng.forced is not None and not torch.is_tensor(rng.forced):
        rng = Rng.replay_moves(torch.from_numpy(np.ascontiguousarray(rng.forced, np.int8).reshape(n, -1)).to(dev))
    out = eng.selfplay(model1.slot, model2.slot, n, i1, i2, greedy=greedy, precision=model1.precision, rng=rng,
                       rec_cap=rec_cap, want_moves=want_moves)
    torch.cuda.synchronize(dev)
    res = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    res["final"] = boards.from_bitboards(res["final_p1"].view(np.uint64), res["final_p2"].view(np.uint64))
    return res


def swapped_states(rec_own, rec_opp, n_rec):
    """The reference records state*(3-state)^2/2 (colours swapped): learner's stones become 2, opponent's 1."""
    k = int(n_rec)
    return list(boards.from_bitboards(np.asarray(rec_opp[:k]).view(np.uint64), np.asarray(rec_own[:k]).view(np.uint64)))


class Game:
    seed = 0

    def __init__(self, model1, model2, uniforms=None, device=0):
        self.state = boards.start_state()
        self.states = []
        self.actions = []
        self.stone_num = 4
        self.pass_flg = False
        self.model1 = model1
        self.model2 = model2
        self.device = device
        self._uniforms = uniforms
        self._game_id = next(_game_counter)

    def __call__(self):
        if self._uniforms is not None:
            rng = Rng.replay_uniforms(np.asarray(self._uniforms, np.float64).reshape(1, -1))
        else:
            rng = Rng.philox(seed=type(self).seed, game_id0=self._game_id, stream_id=STREAM_SELFPLAY)
        out = play_games(self.model1, self.model2, 1, self.state.reshape(1, 8, 8), rng=rng, want_moves=True, device=self.device)
        self.state = out["final"][0]
        self.states = swapped_states(out["rec_own"][0], out["rec_opp"][0], out["n_rec"][0])
        self.actions = [int(a) for a in out["rec_action"][0][:int(out["n_rec"][0])]]
        self.moves = [int(a) for a in out["moves"][0] if a >= 0]
        self.stone_num = 64
        return self.states, self.actions, self.judge()

    def judge(self):
        myself = int(np.sum(self.state == 1))
        opponent = int(np.sum(self.state == 2))
        return 1 if myself > opponent else (-1 if myself < opponent else 0)
