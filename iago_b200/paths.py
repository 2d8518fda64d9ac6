"""Where the committed-by-the-reference weight files live.

The reference opens them cwd-relative ('./models/sl_model.npz' MCTS.py:83, './models/rollout_model.npz'
mcts_self_play.py:19, '../models/...' under src/).  The drop-in keeps that first, then $IAGO_MODELS, then the
repo-local copy made by oracle/fetch_ref.py (baseline/_ref/models).
"""
import os

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model_path(name: str) -> str:
    cands = [os.path.join(".", "models", name), os.path.join("..", "models", name)]
    if os.environ.get("IAGO_MODELS"):
        cands.insert(0, os.path.join(os.environ["IAGO_MODELS"], name))
    cands.append(os.path.join(_REPO, "baseline", "_ref", "models", name))
    for p in cands:
        if os.path.isfile(p):
            return p
    raise FileNotFoundError(f"weight file {name!r} not found (tried {cands}); set IAGO_MODELS")
