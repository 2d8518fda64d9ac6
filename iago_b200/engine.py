"""Engine: one CUDA context of libiago_b200.so (one per GPU / per process) with tensor-level entry points.

Device API  : torch CUDA tensors in and out (int64 tensors carry the 64-bit bitboards), launches are ordered
              on torch's current stream, nothing synchronises.
Host API    : numpy arrays in and out through the library's *_host entry points (pinned staging, H2D, kernel,
              D2H inside the call) — what the reference-style facade classes use.
"""
import ctypes as C
import os
import dataclasses
from typing import Optional

import numpy as np

from . import _lib
from ._lib import IagoError, IagoRng, check

RNG_PHILOX, RNG_UNIFORMS, RNG_FORCED = 0, 1, 2
STREAM_ROLLOUT, STREAM_SELFPLAY, STREAM_MCTS, STREAM_ENV, STREAM_VALUEGEN = 0, 1, 2, 3, 4


# Inference default of the conv nets: precision 2 = fp16 main product + FP8 (E4M3) cross terms, 2 MMA units per K step; max-abs logit error
# 3e-3 on sl_model.npz, value 5e-5, legal arg-max identical on every harvested position (tests/test_nets_gpu.py; north-star bar: 1e-2 and
# 99.9 %).  Precision 3 (fp16 hi/lo split, 3 MMAs, ~1e-4) is what the trainers' forward passes and the bit-level parity tests use.
DEFAULT_PRECISION = int(os.environ.get("IAGO_DEFAULT_PRECISION", "2"))


def fresh_seed():
    """A seed nobody chose: what `seed=None` means in the facades (the reference draws from np.random / random, which differ from run to
    run; a fixed default would replay the same games every time)."""
    return int.from_bytes(os.urandom(8), "little") >> 1


def _prec(precision):
    return DEFAULT_PRECISION if precision is None else int(precision)


@dataclasses.dataclass
class Rng:
    """How games draw moves (include/iago_b200.h `iago_rng`). One uniform per stone placed, none per pass."""
    mode: int = RNG_PHILOX
    seed: int = 0
    game_id0: int = 0
    stream_id: int = STREAM_ROLLOUT
    uniforms: object = None  # [n, stride] float64 (torch cuda tensor for the device API, ndarray for the host API)
    forced: object = None    # [n, stride] int8

    @staticmethod
    def philox(seed=0, game_id0=0, stream_id=STREAM_ROLLOUT):
        return Rng(RNG_PHILOX, seed, game_id0, stream_id)

    @staticmethod
    def replay_uniforms(uniforms):
        return Rng(RNG_UNIFORMS, uniforms=uniforms)

    @staticmethod
    def replay_moves(forced):
        return Rng(RNG_FORCED, forced=forced)


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _nptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load_library()
        self.device = int(device)
        ctx = C.c_void_p()
        check(self.lib.iago_ctx_create(self.device, C.byref(ctx)))
        self.ctx = ctx
        self._keep = []
        self._slot_owner = {}   # net slot -> owner tag; slots handed out by alloc_slot() (facade models, trainers)

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.iago_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    def load_rollout(self, conv1_W, bias2_b):
        """RolloutPolicy parameters (network.py:49-64): conv1/W (1,2,3,3), bias2/b (64,)."""
        W = np.ascontiguousarray(conv1_W, np.float32).reshape(18)
        b = np.ascontiguousarray(bias2_b, np.float32).reshape(64)
        check(self.lib.iago_load_rollout(self.ctx, _nptr(W), _nptr(b)))

    def load_rollout_npz(self, path):
        with np.load(path) as z:
            pre = "predictor/" if "predictor/conv1/W" in z.files else ""
            self.load_rollout(z[pre + "conv1/W"], z[pre + "bias2/b"])

    N_SLOTS = 32   # IAGO_NET_SLOTS of include/iago_b200.h

    def alloc_slot(self, owner):
        """A free net slot for `owner` (any hashable tag), highest numbers first so that explicit low slot numbers used by scripts
        stay free.  Raises when all slots are taken: a model object never silently overwrites another object's weights."""
        for s in range(self.N_SLOTS - 1, -1, -1):
            if s not in self._slot_owner:
                self._slot_owner[s] = owner
                return s
        raise _lib.IagoError(f"all {self.N_SLOTS} net slots of device {self.device} are in use; close() a model or trainer to free one")

    def free_slot(self, slot, owner=None):
        if slot in self._slot_owner and (owner is None or self._slot_owner[slot] == owner):
            del self._slot_owner[slot]

    def load_net(self, slot, params, kind=None, owner=None):
        """SLPolicy / Value weights into a resident slot. `params`: dict in the reference's npz key layout, or a path.
        A slot handed out by alloc_slot() can only be (re)loaded by its owner."""
        from . import npz
        held = self._slot_owner.get(int(slot))
        if held is not None and held != owner:
            raise _lib.IagoError(f"net slot {slot} belongs to {held!r}; use another slot or alloc_slot()")
        if isinstance(params, (str, bytes)) or hasattr(params, "__fspath__"):
            params = npz.read_npz(params)
        kind = npz.detect_kind(params) if kind is None else kind
        flat = npz.flatten(params, kind)
        check(self.lib.iago_load_net(self.ctx, int(slot), int(kind), _nptr(flat), flat.size))
        return kind

    # ------------------------------------------------------------------ helpers
    def _stream(self, stream):
        if stream is not None:
            return C.c_void_p(int(stream))
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _dev(self):
        return _torch().device("cuda", self.device)

    def _check_i64(self, *ts):
        torch = _torch()
        n = ts[0].numel()
        for t in ts:
            if not (t.is_cuda and t.dtype == torch.int64 and t.is_contiguous() and t.numel() == n):
                raise IagoError("bitboard tensors must be contiguous CUDA int64 tensors of equal length")
        return n

    def _rng_struct(self, rng: Rng, n: int, host: bool):
        r = IagoRng(mode=rng.mode, stream_id=rng.stream_id, seed=rng.seed & (2**64 - 1), game_id0=rng.game_id0,
                    uniforms=None, u_stride=0, forced=None, f_stride=0)
        keep = None
        if rng.mode == RNG_UNIFORMS:
            if host:
                keep = np.ascontiguousarray(rng.uniforms, np.float64).reshape(n, -1)
                r.uniforms, r.u_stride = keep.ctypes.data, keep.shape[1]
            else:
                keep = rng.uniforms.contiguous().view(n, -1)
                assert keep.is_cuda and keep.dtype == _torch().float64
                r.uniforms, r.u_stride = keep.data_ptr(), keep.shape[1]
        elif rng.mode == RNG_FORCED:
            if host:
                keep = np.ascontiguousarray(rng.forced, np.int8).reshape(n, -1)
                r.forced, r.f_stride = keep.ctypes.data, keep.shape[1]
            else:
                keep = rng.forced.contiguous().view(n, -1)
                assert keep.is_cuda and keep.dtype == _torch().int8
                r.forced, r.f_stride = keep.data_ptr(), keep.shape[1]
        return r, keep

    # ------------------------------------------------------------------ device API (torch tensors)
    def legal_actions(self, p1, p2, color, stream=None):
        """Batched GameFunctions.legal_actions: int64 mask per board (bit k = action k legal for `color`)."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        assert color.is_cuda and color.dtype == torch.uint8 and color.numel() == n
        out = torch.empty(n, dtype=torch.int64, device=self._dev())
        check(self.lib.iago_legal_actions(self.ctx, _ptr(p1), _ptr(p2), _ptr(color), _ptr(out), n, self._stream(stream)))
        return out

    def place_stone(self, p1, p2, action, color, stream=None):
        """Batched GameFunctions.place_stone, in place on p1/p2 (action -1 = no-op)."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        assert action.is_cuda and action.dtype == torch.int8 and action.numel() == n
        assert color.is_cuda and color.dtype == torch.uint8 and color.numel() == n
        check(self.lib.iago_place_stone(self.ctx, _ptr(p1), _ptr(p2), _ptr(action), _ptr(color), n, self._stream(stream)))
        return p1, p2

    def rollout_logits(self, p1, p2, color, stream=None):
        torch = _torch()
        n = self._check_i64(p1, p2)
        out = torch.empty((n, 64), dtype=torch.float32, device=self._dev())
        check(self.lib.iago_rollout_logits(self.ctx, _ptr(p1), _ptr(p2), _ptr(color), _ptr(out), n, self._stream(stream)))
        return out

    def policy_forward(self, slot, p1, p2, color, probs=True, precision=None, out=None, stream=None):
        """SLPolicy forward for n bitboard positions -> (n,64) float32 CUDA tensor (probabilities, or logits if probs=False)."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        if out is None:
            out = torch.empty((n, 64), dtype=torch.float32, device=self._dev())
        check(self.lib.iago_policy_forward(self.ctx, int(slot), _ptr(p1), _ptr(p2), _ptr(color), n, _ptr(out),
                                           1 if probs else 0, _prec(precision), self._stream(stream)))
        return out

    def policy_forward_acts(self, slot, p1, p2, color, precision=3, stream=None):
        """SLPolicy forward keeping every block's output: (logits (n,64), [acts_l (n,C_l,8,8) for the 8 blocks]) CUDA tensors."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        dev = self._dev()
        logits = torch.empty((n, 64), dtype=torch.float32, device=dev)
        acts = [torch.empty((n, 64 if l == 0 else 128, 8, 8), dtype=torch.float32, device=dev) for l in range(8)]
        ptrs = (C.c_void_p * 8)(*[a.data_ptr() for a in acts])
        check(self.lib.iago_policy_forward_acts(self.ctx, int(slot), _ptr(p1), _ptr(p2), _ptr(color), n, _ptr(logits),
                                                C.cast(ptrs, C.c_void_p), _prec(precision), self._stream(stream)))
        return logits, acts

    def value_forward_acts(self, slot, p1, p2, color, precision=3, stream=None):
        """Value forward keeping every trunk block's output: (values (n,), [acts_l (n,C_l,8,8) for the 8 blocks]) CUDA tensors."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        dev = self._dev()
        values = torch.empty(n, dtype=torch.float32, device=dev)
        acts = [torch.empty((n, 64 if l == 0 else 128, 8, 8), dtype=torch.float32, device=dev) for l in range(8)]
        ptrs = (C.c_void_p * 8)(*[a.data_ptr() for a in acts])
        check(self.lib.iago_value_forward_acts(self.ctx, int(slot), _ptr(p1), _ptr(p2), _ptr(color), n, _ptr(values),
                                               C.cast(ptrs, C.c_void_p), _prec(precision), self._stream(stream)))
        return values, acts

    def value_forward(self, slot, p1, p2, color, precision=None, out=None, stream=None):
        """Value forward for n bitboard positions -> (n,) float32 CUDA tensor."""
        torch = _torch()
        n = self._check_i64(p1, p2)
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=self._dev())
        check(self.lib.iago_value_forward(self.ctx, int(slot), _ptr(p1), _ptr(p2), _ptr(color), n, _ptr(out),
                                          _prec(precision), self._stream(stream)))
        return out

    def rollout_sample(self, p1, p2, color, rng: Optional[Rng] = None, draw=0, stream=None):
        """One Simulate.get_action draw per board; int8 actions, -1 where the mover must pass."""
        torch = _torch()
        rng = rng or Rng()
        n = self._check_i64(p1, p2)
        out = torch.empty(n, dtype=torch.int8, device=self._dev())
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_rollout_sample(self.ctx, _ptr(p1), _ptr(p2), _ptr(color), n, C.byref(r), int(draw),
                                           _ptr(out), self._stream(stream)))
        return out

    def rollout(self, p1, p2, color, rng: Optional[Rng] = None, want_moves=False, counters=None, out=None, stream=None):
        """n independent Simulate(state)(color) games in one lockstep kernel. Returns a dict of CUDA tensors."""
        torch = _torch()
        rng = rng or Rng()
        n = self._check_i64(p1, p2)
        assert color.is_cuda and color.dtype == torch.uint8 and color.numel() == n
        dev = self._dev()
        if out is None:
            out = dict(result=torch.empty(n, dtype=torch.int8, device=dev),
                       final_p1=torch.empty(n, dtype=torch.int64, device=dev),
                       final_p2=torch.empty(n, dtype=torch.int64, device=dev),
                       n_moves=torch.empty(n, dtype=torch.int32, device=dev),
                       moves=torch.empty((n, 64), dtype=torch.int8, device=dev) if want_moves else None)
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_rollout(self.ctx, _ptr(p1), _ptr(p2), _ptr(color), n, C.byref(r), _ptr(out["result"]),
                                    _ptr(out["final_p1"]), _ptr(out["final_p2"]), _ptr(out.get("n_moves")),
                                    _ptr(out.get("moves")), _ptr(counters), self._stream(stream)))
        return out

    def selfplay(self, slot_learner, slot_opponent, n, init_p1=None, init_p2=None, greedy=False, precision=None,
                 rng: Optional[Rng] = None, rec_cap=40, want_moves=False, stream=None):
        """n lockstep rl_self_play.Game(model1, model2)() games. Returns a dict of CUDA tensors (+ 'stats' host ints)."""
        torch = _torch()
        rng = rng or Rng(stream_id=STREAM_SELFPLAY)
        dev = self._dev()
        if init_p1 is not None:
            self._check_i64(init_p1, init_p2)
            assert init_p1.numel() == n
        out = dict(final_p1=torch.empty(n, dtype=torch.int64, device=dev), final_p2=torch.empty(n, dtype=torch.int64, device=dev),
                   result=torch.empty(n, dtype=torch.int8, device=dev),
                   rec_own=torch.zeros((n, rec_cap), dtype=torch.int64, device=dev),
                   rec_opp=torch.zeros((n, rec_cap), dtype=torch.int64, device=dev),
                   rec_action=torch.full((n, rec_cap), -1, dtype=torch.int8, device=dev),
                   n_rec=torch.empty(n, dtype=torch.int32, device=dev),
                   moves=torch.empty((n, 64), dtype=torch.int8, device=dev) if want_moves else None)
        stats = (C.c_int64 * 3)()
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_selfplay(self.ctx, int(slot_learner), int(slot_opponent), n, _ptr(init_p1), _ptr(init_p2),
                                     1 if greedy else 0, _prec(precision), C.byref(r), _ptr(out["final_p1"]),
                                     _ptr(out["final_p2"]), _ptr(out["result"]), _ptr(out["rec_own"]), _ptr(out["rec_opp"]),
                                     _ptr(out["rec_action"]), _ptr(out["n_rec"]), int(rec_cap), _ptr(out["moves"]),
                                     C.cast(stats, C.c_void_p), self._stream(stream)))
        out["stats"] = dict(turn_pairs=int(stats[0]), forwards=int(stats[1]), positions=int(stats[2]))
        return out

    def value_selfplay(self, slot_sl, slot_rl, stop_num, precision=None, rng: Optional[Rng] = None, stream=None):
        """n lockstep value_self_play.SelfPlay(stop_num[g])() games (stop_num: int32 CUDA tensor). Returns a dict of CUDA tensors:
        rec_own / rec_opp (the recorded position, mover's view), rec_color, rec_action (the random move, -1 = none), result,
        final_p1 / final_p2, draws (+ 'stats' host ints)."""
        torch = _torch()
        rng = rng or Rng(stream_id=STREAM_VALUEGEN)
        dev = self._dev()
        assert stop_num.is_cuda and stop_num.dtype == torch.int32 and stop_num.is_contiguous()
        n = stop_num.numel()
        i64 = lambda: torch.empty(n, dtype=torch.int64, device=dev)
        out = dict(rec_own=i64(), rec_opp=i64(), rec_color=torch.empty(n, dtype=torch.uint8, device=dev),
                   rec_action=torch.empty(n, dtype=torch.int8, device=dev), result=torch.empty(n, dtype=torch.int8, device=dev),
                   final_p1=i64(), final_p2=i64(), draws=torch.empty(n, dtype=torch.int32, device=dev))
        stats = (C.c_int64 * 2)()
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_value_selfplay(self.ctx, int(slot_sl), int(slot_rl), n, _ptr(stop_num), _prec(precision), C.byref(r),
                                           _ptr(out["rec_own"]), _ptr(out["rec_opp"]), _ptr(out["rec_color"]), _ptr(out["rec_action"]),
                                           _ptr(out["result"]), _ptr(out["final_p1"]), _ptr(out["final_p2"]), _ptr(out["draws"]),
                                           C.cast(stats, C.c_void_p), self._stream(stream)))
        out["stats"] = dict(turns=int(stats[0]), forwards=int(stats[1]))
        return out

    def env_step(self, slot_opponent, p1, p2, stone_num, pass_flg, action, draws, rng: Optional[Rng] = None, precision=None,
                 want_errors=True, stream=None):
        """rl_env.GameEnv.step for n environments, in place on the CUDA state tensors (p1/p2 int64, stone_num/draws int32,
        pass_flg uint8); action int8. Returns (done uint8[n], opp_action int8[n], rejection-limit errors)."""
        torch = _torch()
        rng = rng or Rng(stream_id=STREAM_ENV)
        n = self._check_i64(p1, p2)
        dev = self._dev()
        done = torch.empty(n, dtype=torch.uint8, device=dev)
        opp = torch.empty(n, dtype=torch.int8, device=dev)
        err = C.c_int32(0)
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_env_step(self.ctx, int(slot_opponent), _prec(precision), n, _ptr(p1), _ptr(p2), _ptr(stone_num),
                                     _ptr(pass_flg), _ptr(action), C.byref(r), _ptr(draws), _ptr(done), _ptr(opp),
                                     C.cast(C.byref(err), C.c_void_p) if want_errors else None, self._stream(stream)))
        return done, opp, int(err.value)

    def sample_unmasked(self, probs, own, opp, draws, rng: Optional[Rng] = None, stream=None):
        """get_position's sampler (rl_env.py:152-172) on device tensors: int8 actions (-1 no legal move)."""
        torch = _torch()
        rng = rng or Rng(stream_id=STREAM_ENV)
        n = self._check_i64(own, opp)
        out = torch.empty(n, dtype=torch.int8, device=self._dev())
        err = C.c_int32(0)
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_sample_unmasked(self.ctx, _ptr(probs), _ptr(own), _ptr(opp), n, C.byref(r), _ptr(draws), _ptr(out),
                                            C.cast(C.byref(err), C.c_void_p), self._stream(stream)))
        if err.value:
            raise RecursionError("maximum recursion depth exceeded")
        return out

    def sample_masked(self, probs, own, opp, draws, rng: Optional[Rng] = None, stream=None):
        """get_action_auto's sampler (game.py:101-108) on device tensors: one uniform per board, int8 actions (-1 = no legal move)."""
        torch = _torch()
        rng = rng or Rng(stream_id=STREAM_ENV)
        n = self._check_i64(own, opp)
        out = torch.empty(n, dtype=torch.int8, device=self._dev())
        r, keep = self._rng_struct(rng, n, host=False)
        check(self.lib.iago_sample_masked(self.ctx, _ptr(probs), _ptr(own), _ptr(opp), n, C.byref(r), _ptr(draws), _ptr(out),
                                          self._stream(stream)))
        return out

    # ------------------------------------------------------------------ host API (numpy arrays)
    @staticmethod
    def pinned(shape, dtype):
        """numpy array on page-locked, device-mapped host memory (a torch pin_memory tensor keeps it alive).  When every buffer
        of a rollout_host call is such an array the library skips its staging copy (include/iago_b200.h, iago_rollout_host)."""
        torch = _torch()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        return t.numpy()[:nbytes].view(dtype).reshape(shape)

    def rollout_host_buffers(self, n, want_moves=False, pinned=True):
        """(p1, p2, color, out) for rollout_host, page-locked by default."""
        mk = self.pinned if pinned else (lambda shape, dtype: np.empty(shape, dtype))
        out = dict(result=mk(n, np.int8), final_p1=mk(n, np.uint64), final_p2=mk(n, np.uint64), n_moves=mk(n, np.int32),
                   moves=mk((n, 64), np.int8) if want_moves else None, counters=np.zeros(2, np.uint64))
        return mk(n, np.uint64), mk(n, np.uint64), mk(n, np.uint8), out

    def rollout_host(self, p1, p2, color, rng: Optional[Rng] = None, want_moves=False, out=None):
        """Same as rollout() with HOST buffers: H2D + kernel + D2H inside the C call (synchronous).  Pageable arrays go through
        the context's pinned staging buffer; arrays from pinned() / rollout_host_buffers() are used in place."""
        rng = rng or Rng()
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        if out is None:
            out = dict(result=np.empty(n, np.int8), final_p1=np.empty(n, np.uint64), final_p2=np.empty(n, np.uint64),
                       n_moves=np.empty(n, np.int32), moves=np.empty((n, 64), np.int8) if want_moves else None,
                       counters=np.zeros(2, np.uint64))
        r, keep = self._rng_struct(rng, n, host=True)
        check(self.lib.iago_rollout_host(self.ctx, _nptr(p1), _nptr(p2), _nptr(color), n, C.byref(r),
                                         _nptr(out["result"]), _nptr(out["final_p1"]), _nptr(out["final_p2"]),
                                         _nptr(out.get("n_moves")), _nptr(out.get("moves")), _nptr(out.get("counters"))))
        return out

    HOST_LANES = 4

    def rollout_host_submit(self, lane, p1, p2, color, rng: Optional[Rng] = None, out=None):
        """Asynchronous rollout_host on one of HOST_LANES lanes: returns at once; rollout_host_wait(lane) blocks until `out` (from
        rollout_host_buffers) is filled.  Every array must be page-locked (pinned() / rollout_host_buffers()) and must not be
        touched until the wait.  With two or three lanes in flight the copies of one batch overlap the kernel of another."""
        rng = rng or Rng()
        n = p1.shape[0]
        for a, dt in ((p1, np.uint64), (p2, np.uint64), (color, np.uint8)):
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous and a.shape == (n,)):
                raise ValueError("rollout_host_submit takes contiguous page-locked arrays p1, p2 (uint64) and color (uint8) of one length")
        if out is None:
            raise ValueError("rollout_host_submit needs `out` from rollout_host_buffers(n)")
        r, keep = self._rng_struct(rng, n, host=True)
        check(self.lib.iago_rollout_host_submit(self.ctx, int(lane), _nptr(p1), _nptr(p2), _nptr(color), n, C.byref(r),
                                                _nptr(out["result"]), _nptr(out["final_p1"]), _nptr(out["final_p2"]),
                                                _nptr(out.get("n_moves")), _nptr(out.get("moves"))))
        self._lane_out = getattr(self, "_lane_out", {})
        self._lane_out[int(lane)] = (out, p1, p2, color)   # keeps the buffers alive while the lane is in flight
        return lane

    def rollout_host_wait(self, lane):
        """Blocks until the submission on `lane` is complete and returns its `out` dict (counters filled)."""
        held = getattr(self, "_lane_out", {}).pop(int(lane), None)
        out = held[0] if held else None
        check(self.lib.iago_rollout_host_wait(self.ctx, int(lane), _nptr(out.get("counters") if out else None)))
        return out

    def _to_dev(self, a, dtype):
        torch = _torch()
        a = np.array(a, copy=True)  # broadcast views are read-only; torch wants a writable buffer
        return torch.from_numpy(a.view(dtype) if dtype is not None else a).to(self._dev())

    def legal_actions_host(self, p1, p2, color):
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        m = self.legal_actions(self._to_dev(p1, np.int64), self._to_dev(p2, np.int64), self._to_dev(color, None))
        return m.cpu().numpy().view(np.uint64)

    def place_stone_host(self, p1, p2, action, color):
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        action = np.ascontiguousarray(np.broadcast_to(np.asarray(action, np.int8), (n,)))
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        d1, d2 = self._to_dev(p1, np.int64), self._to_dev(p2, np.int64)
        self.place_stone(d1, d2, self._to_dev(action, None), self._to_dev(color, None))
        return d1.cpu().numpy().view(np.uint64), d2.cpu().numpy().view(np.uint64)

    def rollout_sample_host(self, p1, p2, color, rng: Optional[Rng] = None, draw=0):
        torch = _torch()
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        rng = rng or Rng()
        if rng.mode == RNG_UNIFORMS:
            rng = Rng.replay_uniforms(torch.from_numpy(np.ascontiguousarray(rng.uniforms, np.float64).reshape(n, 1)).to(self._dev()))
        out = self.rollout_sample(self._to_dev(p1, np.int64), self._to_dev(p2, np.int64), self._to_dev(color, None),
                                  rng=rng, draw=draw)
        return out.cpu().numpy()

    def _host_boards(self, p1, p2, color):
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        return self._to_dev(p1, np.int64), self._to_dev(p2, np.int64), self._to_dev(color, None)

    def policy_forward_host(self, slot, p1, p2, color, probs=True, precision=None):
        return self.policy_forward(slot, *self._host_boards(p1, p2, color), probs=probs, precision=_prec(precision)).cpu().numpy()

    def value_forward_host(self, slot, p1, p2, color, precision=None):
        return self.value_forward(slot, *self._host_boards(p1, p2, color), precision=_prec(precision)).cpu().numpy()

    def rollout_logits_host(self, p1, p2, color):
        p1 = np.ascontiguousarray(p1, np.uint64).reshape(-1)
        n = p1.shape[0]
        p2 = np.ascontiguousarray(p2, np.uint64).reshape(n)
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (n,)))
        out = self.rollout_logits(self._to_dev(p1, np.int64), self._to_dev(p2, np.int64), self._to_dev(color, None))
        return out.cpu().numpy()

    # ------------------------------------------------------------------ measurement helpers
    def last_kernel_ms(self):
        ms = C.c_float()
        check(self.lib.iago_last_kernel_ms(self.ctx, C.byref(ms)))
        return float(ms.value)

    def measure_int_peak(self, iters=4096):
        v = C.c_double()
        check(self.lib.iago_measure_int_peak(self.ctx, int(iters), C.byref(v)))
        return float(v.value)

    def sync(self):
        check(self.lib.iago_ctx_sync(self.ctx))


_default = {}


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the reference keeps global state too: chainer.config, np.random)."""
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
