"""SelfPlay — drop-in for /root/reference/value_self_play.py:12-162 and the loop of gen_value_data.py:12-19 on the GPU.

    state, result = SelfPlay(stop_num)()          # one game, like the reference
    out = play_games(stop_nums)                   # n lockstep games (numpy dict)
    generate(n, path="./value_data5.txt")         # gen_value_data.main(): appends records in the reference's text format

The SL policy (./models/sl_model.npz) plays until stone_num reaches stop_num, the side to move records the board from its
own view (its stones 2, the other side's 1) and plays one uniformly random legal move, the RL policy
(./models/rl_model.npz) plays the game out; result = +1 / 0 / -1 for the recorded mover, or -1 when it had no legal move.
"""
import numpy as np
import torch

from . import boards, network, paths
from .engine import RNG_UNIFORMS, Rng, STREAM_VALUEGEN, default_engine

_models = {}


def default_models(device=0):
    """value_self_play.py:26-29: model0 = sl_model.npz, model1 = rl_model.npz (loaded once per device)."""
    if device not in _models:
        _models[device] = (network.SLPolicy(device=device).load(paths.model_path("sl_model.npz")),
                           network.SLPolicy(device=device).load(paths.model_path("rl_model.npz")))
    return _models[device]


def play_games(stop_nums, model0=None, model1=None, *, rng=None, device=0):
    """n lockstep games. Returns numpy arrays: state (n,8,8) float32 recorded boards (mover 2, other side 1), result int8[n],
    color uint8[n] (the recorded mover), action int8[n] (the random move, -1 = none), final (n,8,8), draws int32[n]."""
    if model0 is None or model1 is None:
        d0, d1 = default_models(device)
        model0, model1 = model0 or d0, model1 or d1
    eng = default_engine(device)
    dev = torch.device("cuda", device)
    stop = torch.from_numpy(np.ascontiguousarray(stop_nums, np.int32)).to(dev)
    n = stop.numel()
    if rng is not None and rng.mode == RNG_UNIFORMS and not torch.is_tensor(rng.uniforms):
        rng = Rng.replay_uniforms(torch.from_numpy(np.ascontiguousarray(rng.uniforms, np.float64).reshape(n, -1)).to(dev))
    out = eng.value_selfplay(model0.slot, model1.slot, stop, precision=model0.precision, rng=rng)
    torch.cuda.synchronize(dev)
    r = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()}
    u = lambda a: a.view(np.uint64)
    return dict(state=boards.from_bitboards(u(r["rec_opp"]), u(r["rec_own"])), result=r["result"], color=r["rec_color"],
                action=r["rec_action"], final=boards.from_bitboards(u(r["final_p1"]), u(r["final_p2"])), draws=r["draws"],
                stats=r["stats"])


class SelfPlay:
    seed = 0
    _games = 0

    def __init__(self, stop_num, model0=None, model1=None, uniforms=None, device=0):
        self.state = boards.start_state()
        self.stop_num = stop_num
        self.stone_num = 4
        self.pass_flg = False
        self.model0, self.model1, self.device = model0, model1, device
        self._uniforms = uniforms
        self._game_id = SelfPlay._games
        SelfPlay._games += 1

    def __call__(self):
        if self._uniforms is not None:
            rng = Rng.replay_uniforms(np.asarray(self._uniforms, np.float64).reshape(1, -1))
        else:
            rng = Rng.philox(seed=type(self).seed, game_id0=self._game_id, stream_id=STREAM_VALUEGEN)
        out = play_games([self.stop_num], self.model0, self.model1, rng=rng, device=self.device)
        self.state = out["final"][0]
        self.stone_num = 64
        return out["state"][0], int(out["result"][0])


def generate(size, path="./value_data5.txt", batch=16384, seed=None, device=0, model0=None, model1=None):
    """gen_value_data.main(): `size` games with stop_num ~ randint(4, 64), appended to `path` record by record in the
    reference's format ("\\n" + str(state) + ", \\r", "\\n", "\\n" + str(result) + ", \\r").  Returns (states, results).
    seed=None (default) draws a fresh seed, so that a second call appends NEW games to the file as the reference does; pass a seed to
    reproduce a run."""
    from .engine import fresh_seed
    if seed is None:
        seed = fresh_seed()
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    states, results = [], []
    done = 0
    while done < size:
        n = min(batch, size - done)
        out = play_games(rs.randint(4, 64, size=n), model0, model1, device=device,
                         rng=Rng.philox(seed=seed, game_id0=done, stream_id=STREAM_VALUEGEN))
        states.append(out["state"]); results.append(out["result"])
        if path:
            with open(path, "a") as f:
                for s, r in zip(out["state"], out["result"]):
                    f.write("\n" + str(s) + ", \r")
                    f.write("\n")
                    f.write("\n" + str(int(r)) + ", \r")
        done += n
    return np.concatenate(states), np.concatenate(results)
