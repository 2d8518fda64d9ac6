"""SearchPool: n_trees PV-MCTS searches in lockstep on one GPU (iago_mcts_* of include/iago_b200.h).

This is the batched form a self-play driver uses ("one tree per game, many games per GPU"); `iago_b200.MCTS.MCTS`
is the one-tree facade with the reference's method names on top of it.
"""
import ctypes as C

import numpy as np

from ._lib import IagoMctsParams, check
from .engine import STREAM_MCTS, _prec, default_engine

PASS_INDEX = 64  # visits[:, 64] / q[:, 64] belong to the pass child (action -1, MCTS.py:112-114)


class SearchPool:
    def __init__(self, n_trees=1, max_nodes=65536, max_leaf_batch=256, tree_id0=0, engine=None, device=0):
        self.eng = engine or default_engine(device)
        self.lib = self.eng.lib
        self.n_trees, self.max_nodes, self.max_leaf_batch = int(n_trees), int(max_nodes), int(max_leaf_batch)
        h = C.c_void_p()
        check(self.lib.iago_mcts_create(self.eng.ctx, self.n_trees, self.max_nodes, self.max_leaf_batch, int(tree_id0), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.iago_mcts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return self.eng._stream(None)

    def set_roots(self, p1, p2, color, reset_tree=True):
        T = self.n_trees
        p1 = np.ascontiguousarray(np.broadcast_to(np.asarray(p1, np.uint64), (T,)))
        p2 = np.ascontiguousarray(np.broadcast_to(np.asarray(p2, np.uint64), (T,)))
        color = np.ascontiguousarray(np.broadcast_to(np.asarray(color, np.uint8), (T,)))
        check(self.lib.iago_mcts_set_roots(self.h, p1.ctypes.data, p2.ctypes.data, color.ctypes.data, 1 if reset_tree else 0,
                                           self._stream()))

    def get_roots(self):
        T = self.n_trees
        p1, p2 = np.empty(T, np.uint64), np.empty(T, np.uint64)
        color, done = np.empty(T, np.uint8), np.empty(T, np.int64)
        check(self.lib.iago_mcts_get_roots(self.h, p1.ctypes.data, p2.ctypes.data, color.ctypes.data, done.ctypes.data, self._stream()))
        return p1, p2, color, done

    def search(self, n_playouts, *, slot_policy, slot_value, lmbda=0.5, c_puct=1.0, n_thr=15, leaf_batch=1, virtual_loss=1.0,
               precision=None, cache_value=True, seed=0, forced_v=None, forced_z=None):
        """n_playouts more MCTS.playout calls on every tree. forced_v / forced_z ([n_trees, stride], indexed by the playout
        number since the tree was created) replay given leaf evaluations instead of running the nets (test hook)."""
        p = IagoMctsParams(lmbda=float(lmbda), c_puct=float(c_puct), virtual_loss=float(virtual_loss), n_thr=int(n_thr),
                           leaf_batch=int(leaf_batch), n_playouts=int(n_playouts), slot_policy=int(slot_policy),
                           slot_value=int(slot_value), precision=_prec(precision), cache_value=1 if cache_value else 0,
                           reserved=0, seed=int(seed) & (2**64 - 1), forced_v=None, forced_z=None, forced_stride=0)
        keep = []
        if forced_v is not None:
            fv = np.ascontiguousarray(forced_v, np.float32).reshape(self.n_trees, -1)
            p.forced_v, p.forced_stride = fv.ctypes.data, fv.shape[1]
            keep.append(fv)
        if forced_z is not None:
            fz = np.ascontiguousarray(forced_z, np.int8).reshape(self.n_trees, -1)
            assert p.forced_stride in (0, fz.shape[1])
            p.forced_z, p.forced_stride = fz.ctypes.data, fz.shape[1]
            keep.append(fz)
        check(self.lib.iago_mcts_search(self.h, C.byref(p), self._stream()))

    def root_stats(self):
        """(visits int32[T,65], q float32[T,65], best int8[T]); best = -2 where the root has no children yet."""
        T = self.n_trees
        visits, q, best = np.empty((T, 65), np.int32), np.empty((T, 65), np.float32), np.empty(T, np.int8)
        check(self.lib.iago_mcts_root_stats(self.h, visits.ctypes.data, q.ctypes.data, best.ctypes.data, self._stream()))
        return visits, q, best

    def advance(self, actions, mask=None):
        T = self.n_trees
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(actions, np.int8), (T,)))
        m = None if mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(mask, np.uint8), (T,)))
        check(self.lib.iago_mcts_advance(self.h, a.ctypes.data, None if m is None else m.ctypes.data, self._stream()))

    def export_tree(self, tree=0):
        """Nodes of one tree in pool order: dict of parent, action, n, Q, P, first_child, n_children."""
        cap = self.max_nodes
        out = dict(parent=np.empty(cap, np.int32), action=np.empty(cap, np.int8), n=np.empty(cap, np.int32),
                   Q=np.empty(cap, np.float64), P=np.empty(cap, np.float64), first_child=np.empty(cap, np.int32),
                   n_children=np.empty(cap, np.int32))
        cnt = C.c_int32()
        check(self.lib.iago_mcts_export_tree(self.h, int(tree), cap, out["parent"].ctypes.data, out["action"].ctypes.data,
                                             out["n"].ctypes.data, out["Q"].ctypes.data, out["P"].ctypes.data,
                                             out["first_child"].ctypes.data, out["n_children"].ctypes.data, C.byref(cnt),
                                             self._stream()))
        return {k: v[:cnt.value].copy() for k, v in out.items()}

    def overflows(self):
        c = C.c_int64()
        check(self.lib.iago_mcts_overflows(self.h, C.byref(c)))
        return int(c.value)


def flatten_bfs(tree):
    """Pool-order dump -> breadth-first arrays (parent, action, n, Q, P), children in ascending action order: the canonical
    form the oracle's trees are compared in."""
    order, parent = [0], [-1]
    i = 0
    while i < len(order):
        nd = order[i]
        fc, nc = int(tree["first_child"][nd]), int(tree["n_children"][nd])
        for j in range(nc):
            order.append(fc + j)
            parent.append(i)
        i += 1
    idx = np.array(order)
    action = tree["action"][idx].copy()
    action[0] = 0   # the root has no incoming move (a re-rooted child keeps its old one in the pool)
    return dict(parent=np.array(parent, np.int32), action=action, n=tree["n"][idx], Q=tree["Q"][idx], P=tree["P"][idx])
