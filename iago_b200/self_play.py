"""SelfGame — the API of /root/reference/self_play.py:6-75 (two-AI self-play with the unmasked rejection sampler).

The reference class is dead at HEAD (`SelfGame()` omits Game.__init__'s required argument, self_play.py:84 vs game.py:15),
so what is kept is its surface and turn semantics: get_position_self / turn_self / judge_self / show_self and the loop
`while stone_num < 64: turn_self(1); turn_self(2)` (self_play.py:87-89).  Moves are sampled like rl_env.get_position
(self_play.py:23-28): p = out - min(out) over all 64 cells, re-drawn until legal; both sides run on the GPU through
iago_policy_forward + iago_sample_unmasked.
Deviation, stated: the reference's colour-1 branch overwrites self.state with the colour-swapped board and never swaps it
back (self_play.py:9-12) — a latent bug that would corrupt the game; here the swap is applied to the network input only.
"""
import numpy as np
import torch

from . import boards
from .engine import Rng, STREAM_ENV, default_engine
from .rl_env import _net


class SelfGame:
    def __init__(self, model1, model2, seed=0, device=0, verbose=False):
        self.state = boards.start_state()
        self.stone_num, self.pass_flg, self.play_num = 4, False, 1
        self.model1, self.model2 = model1, model2
        self.gamelog = ""
        self.device, self.verbose, self.seed = device, verbose, seed
        self._draws = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device))

    def valid_pos(self, color):
        p1, p2 = boards.to_bitboards(self.state)
        m = default_engine(self.device).legal_actions_host(p1, p2, color)[0]
        return [[a // 8 + 1, a % 8 + 1] for a in boards.mask_to_actions(m)]

    def place_stone(self, position, color):
        a = (position[0] - 1) * 8 + (position[1] - 1)
        p1, p2 = boards.to_bitboards(self.state)
        q1, q2 = default_engine(self.device).place_stone_host(p1, p2, a, color)
        self.state = boards.from_bitboards(q1, q2)[0]

    def get_position_self(self, color, positions):
        """One move of `color` by its model (self_play.py:8-30): SLPolicy forward + the unmasked rejection sampler, on the GPU."""
        eng = default_engine(self.device)
        p1, p2 = boards.to_bitboards(self.state)
        dev = self._draws.device
        t1 = torch.from_numpy(p1.view(np.int64).copy()).to(dev)
        t2 = torch.from_numpy(p2.view(np.int64).copy()).to(dev)
        col = torch.full((1,), color, dtype=torch.uint8, device=dev)
        net = _net(self.model1 if color == 1 else self.model2)
        probs = eng.policy_forward(net.slot, t1, t2, col, probs=True, precision=net.precision)
        own, opp = (t1, t2) if color == 1 else (t2, t1)
        a = int(eng.sample_unmasked(probs, own, opp, self._draws, Rng.philox(seed=self.seed, stream_id=STREAM_ENV))[0])
        return [a // 8 + 1, a % 8 + 1]

    def turn_self(self, color):
        players = ["AI1", "AI2"]
        positions = self.valid_pos(color)
        if self.verbose:
            print("Valid choice:", positions)
        if len(positions) > 0:
            position = self.get_position_self(color, positions)
            self.place_stone(position, color)
            if self.verbose:
                self.show_self()
            self.pass_flg = False
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": " + str(position) + "\n"
            self.stone_num += 1
        else:
            if self.pass_flg:
                self.stone_num = 64
            if self.verbose:
                print(players[color - 1] + " pass.")
            self.pass_flg = True
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": Pass\n"
        self.play_num += 1

    def show_self(self):
        print("   1   2   3   4   5   6   7   8   ")
        for i in range(8):
            print(" " + "-" * 34)
            print(str(i + 1) + "|" + "|".join({0: "   ", 1: " X ", 2: " O "}[int(v)] for v in self.state[i]) + "|")
        print(" " + "-" * 33)
        print("X(AI1):" + str(int(np.sum(self.state == 1))) + ", O(AI2):" + str(int(np.sum(self.state == 2)))
              + ", Empty:" + str(int(np.sum(self.state == 0))))

    def judge_self(self):
        ai1, ai2 = int(np.sum(self.state == 1)), int(np.sum(self.state == 2))
        if self.verbose:
            print("AI1 WIN!" if ai1 > ai2 else ("AI1 LOSE" if ai1 < ai2 else "DRAW"))
        return "X(AI1):" + str(ai1) + ", O(AI2):" + str(ai2) + ", Empty:" + str(int(np.sum(self.state == 0)))

    def __call__(self):
        while self.stone_num < 64:   # self_play.py:87-89
            self.turn_self(1)
            self.turn_self(2)
        return self.judge_self()
