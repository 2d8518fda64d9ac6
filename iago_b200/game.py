"""GameFunctions — drop-in for /root/reference/game.py:153-235, computed by the CUDA rules kernels.

Same names, argument order and return types as the reference:
  legal_actions(state, color) -> ascending list[int]          game.py:209-235
  place_stone(state, action, color) -> state (mutated in place; -1 = pass = no-op; no legality check)  game.py:179-207
  make_state_var(state, color) -> float32 (1,2,8,8), channel 0 = opponent, channel 1 = mover          game.py:167-174
  ac2pos(actions), is_outside(pos)                             game.py:155-165
plus *_batch variants on bitboards, which is what a batched caller should use (one launch for N boards).
The single-board calls go to the GPU too (N = 1): there is no CPU rules implementation in this package.
"""
import numpy as np

from . import boards
from .engine import default_engine


class GameFunctions:
    device = 0

    @classmethod
    def ac2pos(cls, actions):
        return [[a // 8 + 1, a % 8 + 1] for a in actions]

    @classmethod
    def is_outside(cls, pos):
        return pos[0] < 0 or pos[0] > 7 or pos[1] < 0 or pos[1] > 7

    @classmethod
    def make_state_var(cls, state, color):
        """Input planes of the networks. Returns a plain float32 ndarray (the reference wraps it in chainer.Variable)."""
        s = np.asarray(state)
        mover = (s == color)
        opp = (s == 3 - color)
        return np.stack([opp, mover], axis=0).astype(np.float32).reshape(1, 2, 8, 8)

    @classmethod
    def legal_actions_batch(cls, p1, p2, color):
        """uint64 legal masks for N boards (numpy in/out)."""
        return default_engine(cls.device).legal_actions_host(p1, p2, color)

    @classmethod
    def place_stone_batch(cls, p1, p2, action, color):
        return default_engine(cls.device).place_stone_host(p1, p2, action, color)

    @classmethod
    def legal_actions(cls, state, color):
        p1, p2 = boards.to_bitboards(state)
        return boards.mask_to_actions(cls.legal_actions_batch(p1, p2, color)[0])

    @classmethod
    def place_stone(cls, state, action, color):
        if action == -1:
            return state
        p1, p2 = boards.to_bitboards(state)
        q1, q2 = cls.place_stone_batch(p1, p2, action, color)
        state[...] = boards.from_bitboards(q1, q2, dtype=state.dtype)[0]
        return state


class Game:
    """Drop-in for game.Game (game.py:13-151): human (or SLPolicy when auto) as colour 1 against PV-MCTS as colour 2.

    Same attributes and turn semantics as the reference: `turn(color, auto)` (game.py:115-142) with update_with_move on
    both sides' moves and on passes (-1), the `stone_num > 62` shortcut of get_action_auto (game.py:97-98), gamelog format
    and save_gamelog.  Keyword extras are forwarded to iago_b200.MCTS.MCTS (n_playouts, leaf_batch, ...)."""

    def __init__(self, auto, seed=None, verbose=True, **mcts_kw):
        import torch
        from datetime import datetime
        from . import network
        from .MCTS import MCTS
        from .paths import model_path
        if auto:
            self.p1 = "IaGo(SLPolicy)"
            self.model = network.SLPolicy().load(model_path("sl_model.npz"))
        else:
            self.p1 = "You"
            self.model = None
        self.p2 = "IaGo(PV-MCTS)"
        self.state = boards.start_state()
        self.stone_num, self.play_num, self.pass_flg = 4, 1, False
        self.date = datetime.now().strftime("%Y-%m-%d-%H-%M")
        self.gamelog = "IaGo \n" + self.date + "\n"
        self.mcts = MCTS(**mcts_kw)
        from .engine import fresh_seed
        self.verbose, self.seed = verbose, (fresh_seed() if seed is None else seed)   # None: a different game every run, as with np.random
        self._draws = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", GameFunctions.device))

    def show(self):
        print("   1   2   3   4   5   6   7   8   ")
        for i in range(8):
            print(" " + "-" * 33)
            print(str(i + 1) + "|" + "|".join({0: "   ", 1: " X ", 2: " O "}[int(v)] for v in self.state[i]) + "|")
        print(" " + "-" * 33)
        print(self.p1 + "(X):" + str(int(np.sum(self.state == 1))) + ", " + self.p2 + "(O):" + str(int(np.sum(self.state == 2)))
              + ", Empty:" + str(int(np.sum(self.state == 0))))
        print("\n")

    def judge(self):
        p1, p2 = int(np.sum(self.state == 1)), int(np.sum(self.state == 2))
        if self.verbose:
            print((self.p1 + " WIN!") if p1 > p2 else ((self.p2 + " WIN") if p1 < p2 else "DRAW"))
        return self.p1 + ":" + str(p1) + ", " + self.p2 + ":" + str(p2) + ", Empty:" + str(int(np.sum(self.state == 0)))

    def safeinput(self):
        import re
        while True:
            line = input()
            if re.fullmatch(r"\d[,]\d", line):
                return line.split(",")
            print("Try again.")

    def get_action(self, color, actions):
        if color == 1:
            while True:
                print("Your turn. Choose a position!")
                position = [int(e) for e in self.safeinput()]
                action = (position[0] - 1) * 8 + (position[1] - 1)
                if action in actions:
                    break
                print("This position is invalid. Choose another position")
            self.mcts.update_with_move(action)
        else:
            if self.verbose:
                print("Thinking... Wait a second.")
            action = self.mcts.get_move(self.state, 2)
            self.mcts.update_with_move(action)
        return action

    def get_action_auto(self, color, actions):
        import torch
        from .engine import Rng, STREAM_ENV
        if self.stone_num > 62 and len(actions) == 1:
            return actions[0]                      # game.py:97-98 (the trees are not advanced on this path either)
        if color == 1:
            eng = default_engine(GameFunctions.device)
            p1, p2 = boards.to_bitboards(self.state)
            dev = self._draws.device
            t1 = torch.from_numpy(p1.view(np.int64).copy()).to(dev)
            t2 = torch.from_numpy(p2.view(np.int64).copy()).to(dev)
            col = torch.ones(1, dtype=torch.uint8, device=dev)
            probs = eng.policy_forward(self.model.slot, t1, t2, col, probs=True, precision=self.model.precision)
            action = int(eng.sample_masked(probs, t1, t2, self._draws, Rng.philox(seed=self.seed, stream_id=STREAM_ENV))[0])
            self.mcts.update_with_move(action)
        else:
            action = self.mcts.get_move(self.state, 2)
            self.mcts.update_with_move(action)
        return action

    def turn(self, color, auto):
        players = [self.p1, self.p2]
        actions = GameFunctions.legal_actions(self.state, color)
        if self.verbose:
            print("Valid choice:", GameFunctions.ac2pos(actions))
        if len(actions) > 0:
            action = self.get_action_auto(color, actions) if auto else self.get_action(color, actions)
            position = [action // 8 + 1, action % 8 + 1]
            if self.verbose:
                print(position)
            self.state = GameFunctions.place_stone(self.state, action, color)
            self.stone_num += 1
            if self.verbose:
                self.show()
            self.pass_flg = False
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": " + str(position) + "\n"
        else:
            if self.pass_flg:
                self.stone_num = 64
            if self.verbose:
                print(players[color - 1] + " pass.")
            self.pass_flg = True
            self.mcts.update_with_move(-1)
            self.gamelog += "[" + str(self.play_num) + "]" + players[color - 1] + ": Pass\n"
        self.play_num += 1

    def save_gamelog(self):
        import os
        filename = "./gamelog/" + self.date + ".txt"
        os.makedirs(os.path.dirname(filename), exist_ok=True)
        with open(filename, "w") as f:
            f.write(self.gamelog)


def main():
    import argparse
    parser = argparse.ArgumentParser(description="IaGo:")
    parser.add_argument("--auto", "-a", type=bool, default=False, help="Set True for auto play between MCTS and SLPolicy")
    parser.add_argument("--playouts", type=int, default=None, help="fixed playouts per move instead of the 10 s budget")
    args = parser.parse_args()
    print("\n" + "*" * 34 + "\n" + "*" * 11 + "Game Start!!" + "*" * 11 + "\n" + "*" * 34 + "\n")
    game = Game(args.auto, n_playouts=args.playouts, leaf_batch=64 if args.playouts else 1)
    game.show()
    while game.stone_num < 64:
        game.turn(1, args.auto)
        game.turn(2, args.auto)
    print("\n" + "*" * 34 + "\n" + "*" * 12 + "Game End!!" + "*" * 12 + "\n" + "*" * 34)
    jd = game.judge()
    print(jd)
    game.gamelog += jd + "\n"
    game.save_gamelog()


if __name__ == "__main__":
    main()
