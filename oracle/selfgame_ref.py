"""numpy restatement of self_play.SelfGame (self_play.py:6-89) — TEST INFRASTRUCTURE ONLY.

get_position_self (self_play.py:8-30): the mover's SLPolicy output on the side-to-move-normalised board, p = out - min(out) over
all 64 cells, np.random.choice, the whole call repeated (new draw) until the cell is legal — the same sampler as
rl_env.get_position, restated once in oracle/env_ref.choice_unmasked and pinned here to the UNMODIFIED function through
tests/golden/selfgame.npz (239 calls made unbound on a stand-in object, oracle/gen_golden.py gen_selfgame).
turn_self / the main loop (self_play.py:46-64, 87-89): legal moves -> move or pass; two consecutive passes set stone_num = 64;
`while stone_num < 64: turn_self(1); turn_self(2)`; gamelog line per turn.
Deviation shared with the product, stated: the reference's colour-1 branch overwrites self.state with the colour-swapped board
and never swaps it back (self_play.py:9-12); here, as in iago_b200/self_play.py, the swap is applied to the network input only.
Rules come from the C oracle; probabilities are injected (policy_func(state, color) -> float32[64]).
"""
import numpy as np

from . import cref
from .env_ref import choice_unmasked


class RefSelfGame:
    def __init__(self, policy_func, uniforms):
        self.policy_func, self.uniforms = policy_func, np.asarray(uniforms, np.float64)
        self.state = cref.start_board()
        self.stone_num, self.pass_flg, self.play_num, self.draws = 4, False, 1, 0
        self.gamelog, self.moves = "", []

    def get_position_self(self, color, legal):
        prob = self.policy_func(self.state, color)
        while True:
            idx = choice_unmasked(prob, self.uniforms[self.draws])
            self.draws += 1
            if idx in legal:
                return idx

    def turn_self(self, color):
        players = ["AI1", "AI2"]
        legal = cref.legal_actions(self.state, color)
        if legal:
            a = self.get_position_self(color, legal)
            cref.place_stone(self.state, a, color)
            self.pass_flg = False
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": " + str([a // 8 + 1, a % 8 + 1]) + "\n"
            self.stone_num += 1
            self.moves.append(a)
        else:
            if self.pass_flg:
                self.stone_num = 64
            self.pass_flg = True
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": Pass\n"
            self.moves.append(-1)
        self.play_num += 1

    def __call__(self):
        while self.stone_num < 64:
            self.turn_self(1)
            self.turn_self(2)
        a, b = int((self.state == 1).sum()), int((self.state == 2).sum())
        return "X(AI1):" + str(a) + ", O(AI2):" + str(b) + ", Empty:" + str(int((self.state == 0).sum()))
