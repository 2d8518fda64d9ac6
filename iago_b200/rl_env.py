"""GameEnv — drop-in for /root/reference/rl_env.py:10-185 (gym-style wrapper of the learner-vs-opponent game).

    env = GameEnv(model1, model2)            # model2 (or model2.predictor) = opponent SLPolicy, colour 2
    obs = env.reset()                        # (1,2,8,8) float32: planes [state==1, state==2]      rl_env.py:26-39
    obs, reward, done, info = env.step(a)    # learner (colour 1) plays a, opponent answers          rl_env.py:41-74
    z = env()                                # judge() from colour 1's view                           rl_env.py:78-79

`env.state` is a plain (8,8) float32 array that callers may poke (reinforce.py:48 does).  Every rule / network /
sampling step runs on the GPU (iago_env_step, iago_legal_actions, iago_place_stone, iago_policy_forward).
`VecGameEnv(n, model2)` is the batched form on device tensors.

Randomness: the reference draws the opponent's moves from the global np.random (one uniform per attempt, re-drawn until
legal, rl_env.py:167-171) and the learner's illegal-move fallback from Python's `random` (rl_env.py:46-48).  Here both come
from one explicit stream per environment (Philox keyed by seed / env id, or `uniforms=` to replay np.random's draws);
the fallback takes positions[floor(u * len)].  Like the reference's global generators, the stream keeps advancing across
episodes: reset() does NOT rewind `draws` (a gym-style loop that resets per episode would otherwise face the same opponent
draws every episode); in replay mode (`uniforms=`) the supplied stream is likewise consumed across resets and must be long
enough for all episodes (the kernel bound-checks it).  `reseed(seed, env_id0)` starts a fresh stream explicitly.
"""
import itertools

import numpy as np
import torch

from . import boards
from .engine import Rng, STREAM_ENV, default_engine

_env_counter = itertools.count()


def _net(model):
    return getattr(model, "predictor", model)   # the reference wraps the nets in L.Classifier (rl_env.py:162-164)


def _obs(state):
    X = np.stack([state == 1, state == 2], axis=0).astype(np.float32)
    return X.reshape(2, 1, 8, 8).transpose(1, 0, 2, 3)


class VecGameEnv:
    """n environments in lockstep on one GPU. State lives in CUDA tensors; `step(actions)` is one iago_env_step."""

    def __init__(self, n, model2, seed=0, env_id0=0, device=0, uniforms=None):
        self.n, self.device = int(n), device
        self.eng = default_engine(device)
        self.model2 = _net(model2)
        self.seed, self.env_id0, self._uniforms = seed, env_id0, uniforms
        self.draws = torch.zeros(self.n, dtype=torch.int32, device=torch.device("cuda", device))   # position in each environment's stream
        self.reset()

    def reseed(self, seed, env_id0=None):
        """A fresh Philox stream (draw index back to 0)."""
        self.seed = seed
        if env_id0 is not None:
            self.env_id0 = env_id0
        self.draws.zero_()

    def _rng(self):
        if self._uniforms is not None:
            u = self._uniforms
            if not torch.is_tensor(u):
                u = torch.from_numpy(np.ascontiguousarray(u, np.float64).reshape(self.n, -1)).to(self.p1.device)
                self._uniforms = u
            return Rng.replay_uniforms(u)
        return Rng.philox(seed=self.seed, game_id0=self.env_id0, stream_id=STREAM_ENV)

    def reset(self):
        dev = torch.device("cuda", self.device)
        self.p1 = torch.full((self.n,), boards.START_P1, dtype=torch.int64, device=dev)
        self.p2 = torch.full((self.n,), boards.START_P2, dtype=torch.int64, device=dev)
        self.stone_num = torch.full((self.n,), 4, dtype=torch.int32, device=dev)
        self.pass_flg = torch.zeros(self.n, dtype=torch.uint8, device=dev)
        return self.p1, self.p2   # `draws` is deliberately left alone: the random stream continues (rl_env.py draws from the global np.random)

    def step(self, actions):
        a = torch.as_tensor(actions, dtype=torch.int8).to(self.p1.device).contiguous()
        done, opp, err = self.eng.env_step(self.model2.slot, self.p1, self.p2, self.stone_num, self.pass_flg, a, self.draws,
                                           rng=self._rng(), precision=self.model2.precision)
        if err:
            raise RecursionError("maximum recursion depth exceeded")   # what the reference's rejection loop dies with
        return (self.p1, self.p2), 0, done, opp

    def judge(self):
        pc = lambda t: torch.tensor([bin(int(v) & (2**64 - 1)).count("1") for v in t.cpu().numpy()])
        a, b = pc(self.p1), pc(self.p2)
        return torch.sign(a - b).to(torch.int8)


class GameEnv:
    seed = 0

    def __init__(self, model1, model2, uniforms=None, device=0):
        self.model1, self.model2 = model1, model2
        self.device = device
        self._uniforms = None if uniforms is None else np.asarray(uniforms, np.float64).reshape(1, -1)
        self._id = next(_env_counter)
        self._vec = VecGameEnv(1, model2, seed=type(self).seed, env_id0=self._id, device=device, uniforms=self._uniforms)
        self.reset()

    def reset(self):
        self.state = boards.start_state()
        self.stone_num = 4
        self.pass_flg = False
        self._vec.reset()
        return _obs(self.state)

    def _push(self):
        p1, p2 = boards.to_bitboards(self.state)
        dev = self._vec.p1.device
        self._vec.p1 = torch.from_numpy(p1.view(np.int64).copy()).to(dev)
        self._vec.p2 = torch.from_numpy(p2.view(np.int64).copy()).to(dev)
        self._vec.stone_num.fill_(int(self.stone_num))
        self._vec.pass_flg.fill_(1 if self.pass_flg else 0)

    def _pull(self):
        self.state = boards.from_bitboards(self._vec.p1.cpu().numpy().view(np.uint64), self._vec.p2.cpu().numpy().view(np.uint64))[0]
        self.stone_num = int(self._vec.stone_num[0])
        self.pass_flg = bool(self._vec.pass_flg[0])

    def step(self, action):
        self._push()   # the caller may have poked self.state (reinforce.py:48)
        _, _, done, opp = self._vec.step(np.array([action], np.int8))
        self._pull()
        self.last_opponent_action = int(opp[0])
        return _obs(self.state), 0, bool(done[0]), None

    def __call__(self):
        return self.judge()

    def is_outside(self, pos):
        return pos[0] < 0 or pos[0] > 7 or pos[1] < 0 or pos[1] > 7

    def place_stone(self, position, color):
        """1-based [row, col] (rl_env.py:88-112)."""
        a = (position[0] - 1) * 8 + (position[1] - 1)
        p1, p2 = boards.to_bitboards(self.state)
        q1, q2 = default_engine(self.device).place_stone_host(p1, p2, a, color)
        self.state = boards.from_bitboards(q1, q2)[0]

    def valid_pos(self, color):
        """1-based [row, col] list in row-major order (rl_env.py:114-138)."""
        p1, p2 = boards.to_bitboards(self.state)
        m = default_engine(self.device).legal_actions_host(p1, p2, color)[0]
        return [[a // 8 + 1, a % 8 + 1] for a in boards.mask_to_actions(m)]

    def judge(self):
        you, ai = int(np.sum(self.state == 1)), int(np.sum(self.state == 2))
        return 1 if you > ai else (-1 if you < ai else 0)
