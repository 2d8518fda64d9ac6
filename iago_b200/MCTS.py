"""MCTS / Node — drop-in for /root/reference/MCTS.py on the GPU-resident search (csrc/mcts.cu).

    mcts = MCTS(lmbda=0.5, c_puct=1, n_thr=15, time_limit=10)      # MCTS.py:80, same defaults
    action = mcts.get_move(state, color)                            # MCTS.py:139-147
    mcts.update_with_move(action)                                   # MCTS.py:149-154   (-1 = pass)
    mcts.root.children[a].n_visits / .Q / .P / .u                   # MCTS.py:10-76 (read-only mirror of the device tree)

Additions (keyword-only): n_playouts (fixed playout budget instead of the wall-clock one), leaf_batch (1 = the
reference's sequential algorithm, >1 = leaf-parallel waves with virtual loss), virtual_loss, seed, device, max_nodes.
The networks load once from the reference's weight files ('./models/sl_model.npz', './models/value_model.npz',
'./models/rollout_model.npz' — MCTS.py:83,85, mcts_self_play.py:19) via iago_b200.paths.model_path.
"""
import time

import numpy as np

from . import boards, network
from .engine import default_engine
from .paths import model_path
from .search import SearchPool, flatten_bfs

_nets = {}


def _load_nets(device):
    if device not in _nets:
        eng = default_engine(device)
        sl = network.SLPolicy(device=device).load(model_path("sl_model.npz"))
        va = network.Value(device=device).load(model_path("value_model.npz"))
        eng.load_rollout_npz(model_path("rollout_model.npz"))
        _nets[device] = (sl, va)
    return _nets[device]


class Node:
    """Read-only mirror of one device node (MCTS.py:10-76 field names)."""

    def __init__(self, parent=None, prob=0):
        self.parent = parent
        self.children = {}
        self.n_visits = 0
        self.Q = 0
        self.u = prob + 0.1
        self.P = prob + 0.1

    def is_root(self):
        return self.parent is None

    def is_leaf(self):
        return len(self.children) < 1

    def U(self, c_puct):
        return c_puct * self.P * np.sqrt(self.parent.n_visits) / (0.01 + self.n_visits)

    def get_value(self):
        return self.Q + self.u


def _mirror(tree, c_puct):
    nodes = []
    for i in range(len(tree["n"])):
        nd = Node()
        nd.n_visits, nd.Q, nd.P = int(tree["n"][i]), float(tree["Q"][i]), float(tree["P"][i])
        nodes.append(nd)
    for i, nd in enumerate(nodes):
        fc, nc = int(tree["first_child"][i]), int(tree["n_children"][i])
        for j in range(nc):
            ch = nodes[fc + j]
            ch.parent = nd
            nd.children[int(tree["action"][fc + j])] = ch
    for nd in nodes:
        nd.u = nd.U(c_puct) if nd.parent is not None else nd.P
    return nodes[0]


class MCTS:
    def __init__(self, lmbda=0.5, c_puct=1, n_thr=15, time_limit=10, *, n_playouts=None, leaf_batch=1, virtual_loss=1.0,
                 seed=None, device=0, max_nodes=1 << 18, precision=None, cache_value=True, root_trees=1):
        self.lmbda, self.c_puct, self.n_thr, self.time_limit = lmbda, c_puct, n_thr, time_limit
        from .engine import fresh_seed
        self.n_playouts, self.leaf_batch, self.virtual_loss = n_playouts, leaf_batch, virtual_loss
        self.seed = fresh_seed() if seed is None else seed   # None: rollouts differ from run to run, like the reference's np.random draws
        self.precision, self.cache_value = precision, cache_value
        self.policy_net, self.value_net = _load_nets(device)
        # root_trees = R > 1: root parallelisation on one GPU — R independent trees on the same root (tree ids 0..R-1: their rollouts
        # draw from different Philox streams), n_playouts / R playouts each in lockstep waves, the move chosen from the SUMMED root
        # visit counts (the single-GPU form of SURVEY.md 8e's optional exchange; parallel.root_parallel_moves is the multi-GPU form).
        # Not the reference's algorithm (one tree): opt-in.  A 16,384-playout move takes 64 dependent waves with one tree, 8 with R = 8.
        self.root_trees = max(1, int(root_trees))
        self.pool = SearchPool(self.root_trees, max_nodes=max_nodes, max_leaf_batch=max(leaf_batch, 1), engine=default_engine(device))
        self._fresh = True
        self.playouts = 0  # playouts run by the last get_move

    # ---- MCTS.py:93-103,135-137: the three evaluators, same names, each one GPU call
    def policy_func(self, state, color, actions):
        p1, p2 = boards.to_bitboards(state)
        prob = self.policy_net.forward_bitboards(p1, p2, color).reshape(64)
        return [(a, prob[a]) for a in actions]

    def value_func(self, state, color):
        p1, p2 = boards.to_bitboards(state)
        return self.value_net.forward_bitboards(p1, p2, color).reshape(1)[0]

    def evaluate_rollout(self, state, color):
        from .mcts_self_play import Simulate
        return Simulate(state)(color)

    def _search(self, n):
        self.pool.search(n, slot_policy=self.policy_net.slot, slot_value=self.value_net.slot, lmbda=self.lmbda,
                         c_puct=self.c_puct, n_thr=self.n_thr, leaf_batch=self.leaf_batch, virtual_loss=self.virtual_loss,
                         precision=self.precision, cache_value=self.cache_value, seed=self.seed)
        self.playouts += n

    def playout(self, state, color, node=None):
        """One MCTS.playout from the root (MCTS.py:105-133). `node` is accepted for signature compatibility; the search
        always starts at the device tree's root."""
        p1, p2 = boards.to_bitboards(state)
        self.pool.set_roots(p1, p2, color, reset_tree=self._fresh)
        self._fresh = False
        self._search(1)

    def get_move(self, state, color):
        p1, p2 = boards.to_bitboards(state)
        self.pool.set_roots(p1, p2, color, reset_tree=self._fresh)
        self._fresh = False
        self.playouts = 0
        if self.n_playouts is not None:
            self._search(-(-int(self.n_playouts) // self.root_trees))
        else:
            start = time.time()
            chunk = max(self.leaf_batch, 1) * 16
            while time.time() - start < self.time_limit:
                self._search(chunk)
        visits, _, best = self.pool.root_stats()
        if (best == -2).all():
            raise ValueError("max() arg is an empty sequence")  # what MCTS.py:147 raises when the root was never expanded
        if self.root_trees == 1:
            return int(best[0])
        v = visits.astype(np.int64).sum(axis=0)              # summed over the trees; index 64 = the pass child
        if v[:64].sum() == 0:
            return -1
        return int(np.argmax(v[:64]))                        # first maximum = lowest action, the reference's tie rule (MCTS.py:147)

    def update_with_move(self, last_move):
        if self._fresh:
            return
        self.pool.advance(np.full(self.root_trees, last_move, np.int8))

    @property
    def root(self):
        if self._fresh:
            return Node(None, 1.0)
        return _mirror(self.pool.export_tree(0), self.c_puct)

    def tree(self):
        """Breadth-first arrays of the current tree (tests)."""
        return flatten_bfs(self.pool.export_tree(0))
