"""Simulate — drop-in for /root/reference/mcts_self_play.py:9-134 on the lockstep rollout kernel.

    Simulate(state)(color) -> +1 / 0 / -1      one rollout-policy game to the end, result for `color`

Differences that are deliberate and documented:
  * the rollout weights are loaded once per process, not re-read from disk per instance (mcts_self_play.py:18-19);
  * uniforms come from a counter-based Philox stream (or a replayed stream), not the global numpy RNG —
    pass `uniforms=` to replay exactly what np.random would have drawn, or set `USE_NUMPY_RNG = True` to make every
    Simulate()(color) consume the global np.random stream exactly as the reference does (one uniform per stone placed, none per
    pass): the unmodified reference MCTS.py then rebuilds its own trees on top of this module (tests/test_dropin_gpu.py);
  * `simulate_batch` runs N games in one launch; the class form is the N = 1 case of the same kernel.
"""
import itertools

import numpy as np

from . import boards
from .engine import Rng, default_engine
from .paths import model_path

USE_NUMPY_RNG = False   # True: Simulate draws from the global np.random like mcts_self_play.py:106 (np.random.choice)

_loaded = {}
_game_counter = itertools.count()


def _engine(device=0):
    eng = default_engine(device)
    if device not in _loaded:
        eng.load_rollout_npz(model_path("rollout_model.npz"))
        _loaded[device] = True
    return eng


def simulate_batch(states=None, colors=1, *, p1=None, p2=None, rng=None, want_moves=False, device=0):
    """N independent Simulate(state)(color). Give `states` (N,8,8) or bitboards p1/p2. Returns a dict of numpy arrays:
    result int8[N], final_p1/final_p2 uint64[N], n_moves int32[N], moves int8[N,64] (if want_moves), counters."""
    if p1 is None:
        p1, p2 = boards.to_bitboards(states)
    return _engine(device).rollout_host(p1, p2, colors, rng=rng, want_moves=want_moves)


def simulate_stream(batches, *, seed=0, game_id0=0, want_moves=False, device=0, in_flight=3):
    """simulate_batch over an iterable of (p1, p2, colors) bitboard batches with up to `in_flight` of them on the GPU at a time
    (Engine.rollout_host_submit / _wait: the copies of one batch overlap the kernel of another).  Yields one result dict per batch, in
    order, as copies; game ids run on from `game_id0` across the batches, so the games are those of one simulate_batch call over the
    concatenation with Rng.philox(seed, game_id0)."""
    eng = _engine(device)
    lanes = max(1, min(int(in_flight), eng.HOST_LANES))
    bufs = [None] * lanes
    pending = []                     # (lane, n) in submission order

    def collect():
        lane, n = pending.pop(0)
        out = eng.rollout_host_wait(lane)
        return {k: (np.array(v[:n]) if k != "counters" else np.array(v)) for k, v in out.items() if v is not None}

    gid = int(game_id0)
    for i, (p1, p2, colors) in enumerate(batches):
        lane = i % lanes
        if len(pending) == lanes:
            yield collect()
        p1 = np.asarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        if bufs[lane] is None or bufs[lane][0].shape[0] < n:
            bufs[lane] = eng.rollout_host_buffers(max(n, 1), want_moves=want_moves)
        b1, b2, bc, out = bufs[lane]
        b1[:n], b2[:n], bc[:n] = p1, np.asarray(p2, np.uint64).reshape(n), np.broadcast_to(np.asarray(colors, np.uint8), (n,))
        view = {k: (v[:n] if (v is not None and k != "counters") else v) for k, v in out.items()}
        eng.rollout_host_submit(lane, b1[:n], b2[:n], bc[:n], rng=Rng.philox(seed=seed, game_id0=gid), out=view)
        pending.append((lane, n))
        gid += n
    while pending:
        yield collect()


class Simulate:
    seed = 0  # class-level Philox key; each instance takes the next game id

    def __init__(self, state, uniforms=None, device=0):
        self.state = np.array(state, dtype=np.float32, copy=True)  # copy.deepcopy(state)
        self.stone_num = 64 - int(np.sum(self.state == 0))
        self._start_empty = int(np.sum(self.state == 0))
        self.pass_flg = False
        self.device = device
        self._uniforms = uniforms
        self._game_id = next(_game_counter)
        _engine(device)

    def _rng(self):
        if self._uniforms is not None:
            return Rng.replay_uniforms(np.asarray(self._uniforms, np.float64).reshape(1, -1))
        return Rng.philox(seed=type(self).seed, game_id0=self._game_id)

    def __call__(self, color):
        rng, np_state = self._rng(), None
        if self._uniforms is None and USE_NUMPY_RNG:
            np_state = np.random.get_state()
            rng = Rng.replay_uniforms(np.random.random_sample(64).reshape(1, -1))
        out = simulate_batch(self.state.reshape(1, 8, 8), color, rng=rng, want_moves=True, device=self.device)
        if np_state is not None:     # leave the global stream where the reference would: one draw per stone placed
            np.random.set_state(np_state)
            if int(out["n_moves"][0]) > 0:
                np.random.random_sample(int(out["n_moves"][0]))
        self.state = boards.from_bitboards(out["final_p1"], out["final_p2"])[0]
        self.moves = [int(a) for a in out["moves"][0] if a >= 0]
        if self.stone_num < 64:
            self.stone_num = 64
        return int(out["result"][0])

    # --- the reference's per-step methods, same names (mcts_self_play.py:36-134), each a GPU call ---
    def legal_actions(self, color):
        p1, p2 = boards.to_bitboards(self.state)
        return boards.mask_to_actions(_engine(self.device).legal_actions_host(p1, p2, color)[0])

    def place_stone(self, state, action, color):
        p1, p2 = boards.to_bitboards(state)
        q1, q2 = _engine(self.device).place_stone_host(p1, p2, action, color)
        state[...] = boards.from_bitboards(q1, q2, dtype=state.dtype)[0]
        return state

    def get_action(self, color, actions):
        """mcts_self_play.py:100-110 — one masked, renormalised rollout-policy draw (k-th stone -> k-th uniform)."""
        p1, p2 = boards.to_bitboards(self.state)
        k = int(np.sum(self.state != 0)) - (64 - self._start_empty)
        if self._uniforms is not None:
            rng = Rng.replay_uniforms(np.asarray(self._uniforms, np.float64).reshape(-1)[k:k + 1])
        else:
            rng = self._rng()
        return int(_engine(self.device).rollout_sample_host(p1, p2, color, rng=rng, draw=k)[0])

    def turn(self, color):
        """mcts_self_play.py:124-134"""
        actions = self.legal_actions(color)
        if len(actions) > 0:
            action = self.get_action(color, actions)
            self.state = self.place_stone(self.state, action, color)
            self.pass_flg = False
            self.stone_num += 1
        else:
            if self.pass_flg:
                self.stone_num = 64
            self.pass_flg = True

    def make_state_var(self, state, color):
        from .game import GameFunctions
        return GameFunctions.make_state_var(state, color)

    def judge(self, color):
        myself = int(np.sum(self.state == color))
        opponent = int(np.sum(self.state == 3 - color))
        return 1 if myself > opponent else (-1 if myself < opponent else 0)
