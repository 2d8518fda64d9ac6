"""Board encodings: the reference's np.float32[8,8] (0 empty / 1 / 2, game.py:26-30) <-> bitboard pairs.

Bit k <-> action k = row*8+col (game.py:184), so ascending bit order is the reference's ascending action list.
Pure data formatting; no rules logic lives here.
"""
import numpy as np

_W = (np.uint64(1) << np.arange(64, dtype=np.uint64))


def start_state():
    """game.py:26-30"""
    s = np.zeros([8, 8], dtype=np.float32)
    s[4, 3] = 1
    s[3, 4] = 1
    s[3, 3] = 2
    s[4, 4] = 2
    return s


START_P1 = (1 << 35) | (1 << 28)
START_P2 = (1 << 27) | (1 << 36)


def to_bitboards(states):
    """(N,8,8) or (8,8) array of {0,1,2} -> (p1, p2) uint64 arrays of shape (N,)."""
    s = np.asarray(states).reshape(-1, 64)
    p1 = ((s == 1).astype(np.uint64) * _W).sum(axis=1, dtype=np.uint64)
    p2 = ((s == 2).astype(np.uint64) * _W).sum(axis=1, dtype=np.uint64)
    return p1, p2


def from_bitboards(p1, p2, dtype=np.float32):
    """(p1, p2) -> (N,8,8) array of {0,1,2}."""
    p1 = np.asarray(p1, np.uint64).reshape(-1, 1)
    p2 = np.asarray(p2, np.uint64).reshape(-1, 1)
    sh = np.arange(64, dtype=np.uint64)
    s = ((p1 >> sh) & np.uint64(1)).astype(dtype) + 2 * ((p2 >> sh) & np.uint64(1)).astype(dtype)
    return s.reshape(-1, 8, 8)


def mask_to_actions(mask):
    """uint64 legal mask -> ascending list of actions (the reference's list order, game.py:209-235)."""
    m = int(mask)
    out = []
    while m:
        low = m & -m
        out.append(low.bit_length() - 1)
        m ^= low
    return out
