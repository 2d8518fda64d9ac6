"""ctypes binding of oracle/libothello_oracle.so (the C restatement in othello_ref.c).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libothello_oracle.so")

RNG_PHILOX, RNG_UNIFORMS, RNG_FORCED = 0, 1, 2

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "othello_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "CC=gcc"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.oracle_legal_actions.argtypes = [_f32p, C.c_int, _i32p]
        L.oracle_legal_actions.restype = C.c_int
        L.oracle_place_stone.argtypes = [_f32p, C.c_int, C.c_int]
        L.oracle_place_stone.restype = None
        L.oracle_perft.argtypes = [_f32p, C.c_int, C.c_int]
        L.oracle_perft.restype = C.c_uint64
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_philox_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
        L.oracle_philox_uniform.restype = C.c_double
        L.oracle_exp32_neg.argtypes = [C.c_float]
        L.oracle_exp32_neg.restype = C.c_float
        L.oracle_rollout_logits.argtypes = [_f32p, C.c_int, _f32p, _f32p, _f32p]
        L.oracle_rollout_logits.restype = None
        L.oracle_policy_is_fast.argtypes = [_f32p, _f32p]
        L.oracle_policy_is_fast.restype = C.c_int
        L.oracle_canon_exp.argtypes = [C.c_double]
        L.oracle_canon_exp.restype = C.c_double
        L.oracle_rollout_sample.argtypes = [_f32p, C.c_int, _f32p, _f32p, C.c_double]
        L.oracle_rollout_sample.restype = C.c_int
        L.oracle_simulate_batch.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int64, _f32p, _f32p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64,
            C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_simulate_batch.restype = C.c_int
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def start_board():
    s = np.zeros((8, 8), np.float32)
    s[4, 3] = s[3, 4] = 1
    s[3, 3] = s[4, 4] = 2
    return s


def legal_actions(state, color):
    out = np.zeros(64, np.int32)
    n = lib().oracle_legal_actions(np.ascontiguousarray(state, np.float32).reshape(64), int(color), out)
    return out[:n].tolist()


def place_stone(state, action, color):
    """In place on a contiguous float32 (8,8) array, like the reference."""
    assert state.dtype == np.float32 and state.flags.c_contiguous
    lib().oracle_place_stone(state.reshape(64), int(action), int(color))
    return state


def perft(state, color, depth):
    return int(lib().oracle_perft(np.ascontiguousarray(state, np.float32).reshape(64), int(color), int(depth)))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return list(o)


def philox_uniform(seed, game, draw, stream=0):
    return float(lib().oracle_philox_uniform(int(seed), int(game), int(draw), int(stream)))


def exp32_neg(x):
    return np.float32(lib().oracle_exp32_neg(float(np.float32(x))))


def rollout_logits(state, color, W, b):
    out = np.zeros(64, np.float32)
    lib().oracle_rollout_logits(np.ascontiguousarray(state, np.float32).reshape(64), int(color),
                                np.ascontiguousarray(W, np.float32).reshape(18),
                                np.ascontiguousarray(b, np.float32).reshape(64), out)
    return out


def policy_is_fast(W, b):
    """Which sampler the weights select (oracle/othello_ref.c header): True = product-of-exponentials tables."""
    W = np.ascontiguousarray(W, np.float32).reshape(18)
    b = np.ascontiguousarray(b, np.float32).reshape(64)
    return bool(lib().oracle_policy_is_fast(W, b))


def canon_exp(x):
    return float(lib().oracle_canon_exp(float(x)))


def rollout_sample(state, color, W, b, u):
    return int(lib().oracle_rollout_sample(np.ascontiguousarray(state, np.float32).reshape(64), int(color),
                                           np.ascontiguousarray(W, np.float32).reshape(18),
                                           np.ascontiguousarray(b, np.float32).reshape(64), float(u)))


def simulate_batch(states, colors, W, b, *, mode=RNG_PHILOX, seed=0, stream=0, game_id0=0,
                   uniforms=None, forced=None, want_moves=True, threads=1):
    """Runs Simulate(state)(color) for every row. Returns dict(final, results, moves, n_moves, n_turns, threads)."""
    states = np.array(states, np.float32, copy=True).reshape(-1, 64)
    n = states.shape[0]
    colors = np.ascontiguousarray(np.broadcast_to(np.asarray(colors, np.int32), (n,)))
    results = np.zeros(n, np.int8)
    moves = np.full((n, 64), -1, np.int8) if want_moves else None
    n_moves = np.zeros(n, np.int32)
    n_turns = np.zeros(n, np.int32)
    up, us, fp, fs = None, 0, None, 0
    if mode == RNG_UNIFORMS:
        uniforms = np.ascontiguousarray(uniforms, np.float64).reshape(n, -1)
        up, us = uniforms.ctypes.data, uniforms.shape[1]
    if mode == RNG_FORCED:
        forced = np.ascontiguousarray(forced, np.int8).reshape(n, -1)
        fp, fs = forced.ctypes.data, forced.shape[1]
    used = lib().oracle_simulate_batch(
        states.ctypes.data, colors.ctypes.data, n,
        np.ascontiguousarray(W, np.float32).reshape(18), np.ascontiguousarray(b, np.float32).reshape(64),
        mode, int(seed), int(stream), int(game_id0), up, us, fp, fs,
        results.ctypes.data, moves.ctypes.data if want_moves else None,
        n_moves.ctypes.data, n_turns.ctypes.data, int(threads))
    return dict(final=states.reshape(n, 8, 8), results=results, moves=moves, n_moves=n_moves,
                n_turns=n_turns, threads=used)


# ---- helpers shared by tests: float board <-> bitboards (bit k <-> action k = row*8+col) ----

def to_bitboards(states):
    s = np.asarray(states, np.float32).reshape(-1, 64)
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    p1 = ((s == 1).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    p2 = ((s == 2).astype(np.uint64) * w).sum(axis=1, dtype=np.uint64)
    return p1, p2


def from_bitboards(p1, p2):
    p1 = np.asarray(p1, np.uint64).reshape(-1, 1)
    p2 = np.asarray(p2, np.uint64).reshape(-1, 1)
    sh = np.arange(64, dtype=np.uint64)
    s = ((p1 >> sh) & np.uint64(1)).astype(np.float32) + 2 * ((p2 >> sh) & np.uint64(1)).astype(np.float32)
    return s.reshape(-1, 8, 8)
