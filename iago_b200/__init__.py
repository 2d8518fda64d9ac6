"""iago_b200 — B200-native batched Othello rollout / self-play / PV-MCTS engine.

Drop-in for the hot path of shionhonda/IaGo (rules, rollout, SL/RL policy and value inference, PV-MCTS,
REINFORCE self-play) behind the reference's Python surface.  All compute runs in hand-written sm_100a CUDA
kernels inside `libiago_b200.so`, reached through a ctypes C ABI (include/iago_b200.h); PyTorch only owns
device memory and streams.  There is no CPU fallback: importing works anywhere, but any compute call raises
if the library or a B200 is missing.
"""
from ._lib import IagoError, load_library  # noqa: F401
from .engine import Engine, Rng, default_engine  # noqa: F401
from . import boards  # noqa: F401

__all__ = ["Engine", "Rng", "default_engine", "IagoError", "load_library", "boards"]
