"""Import the UNMODIFIED reference modules from /root/reference under the chainer stand-in.

TEST INFRASTRUCTURE ONLY (build container; /root/reference does not exist on the
GPU box).  Used by oracle/gen_golden.py and by the optional live-reference tests.

Work-arounds for entry points that are broken at the reference HEAD,
none of which change semantics:
  * NUMBA_DISABLE_JIT=1         — bare @jit on a method fails to type under numba>=0.59
                                   (mcts_self_play.py:36,137)
  * `import MCTS` before `game` — circular import (game.py:10 <-> MCTS.py:8)
  * MCTS.Node.copy = identity   — MCTS.py:106 calls a method Node does not have; identity is
                                   the only reading under which tree statistics persist
  * cwd = reference root        — weight paths are cwd-relative (MCTS.py:83,85 ...)
"""
import importlib
import os
import sys

REF = os.environ.get("IAGO_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REF, "game.py"))


def load():
    """Returns a dict of reference modules: MCTS, game, mcts_self_play, network, rl_self_play."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    os.environ["NUMBA_DISABLE_JIT"] = "1"
    shim = os.path.join(_HERE, "chainer_shim")
    for p in (os.path.join(REF, "src"), REF, shim):
        if p in sys.path:
            sys.path.remove(p)
    # order: shim first, then reference root, then src/ (src/network.py duplicates network.py)
    sys.path[:0] = [shim, REF, os.path.join(REF, "src")]
    os.chdir(REF)
    mods = {}
    mods["network"] = importlib.import_module("network")
    mods["MCTS"] = importlib.import_module("MCTS")  # must precede `game`
    mods["game"] = importlib.import_module("game")
    mods["mcts_self_play"] = importlib.import_module("mcts_self_play")
    mods["rl_self_play"] = importlib.import_module("rl_self_play")
    mods["MCTS"].Node.copy = lambda self: self
    import chainer
    chainer.config.train = False
    chainer.config.enable_backprop = False
    mods["chainer"] = chainer
    return mods
