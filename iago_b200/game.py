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
