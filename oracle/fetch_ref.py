"""Copy what the GPU box needs from the read-only reference tree into the git-ignored baseline/_ref/.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, `gpurun` ships /root/repo only.
baseline/_ref/ is git-ignored (reference material stays out of history) but NOT gpurun-ignored, so the
committed-by-the-reference weight files (models/*.npz — data, not source) travel with the snapshot.
The reference's own .py files are copied next to them ONLY so that bench.py can time the unmodified
Python reference beside the GPU numbers; the product never imports them.

Chainer install outcome (recorded in DESIGN.md): `pip install --no-index --find-links /opt/wheelhouse`
cannot resolve chainer (no wheel, no network); the reference is therefore run under oracle/chainer_shim.
"""
import os
import shutil
import sys

REF = os.environ.get("IAGO_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")

PY = ["game.py", "MCTS.py", "mcts_self_play.py", "network.py", "src/rl_self_play.py"]
MODELS = ["models/rollout_model.npz", "models/sl_model.npz", "models/value_model.npz", "models/rl_model.npz",
          "models/RL/model0.npz", "models/RL/model1.npz", "models/RL/model2.npz"]


def main():
    if not os.path.isdir(REF):
        print(f"fetch_ref: {REF} not present; nothing to do")
        return 0
    for rel in PY + MODELS:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
    print(f"fetch_ref: baseline/_ref populated from {REF}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
