"""CPU restatement of the reference's PV-MCTS (MCTS.py:10-154) — TEST INFRASTRUCTURE ONLY.

Two things live here, both plain Python over numpy scalars so that every arithmetic type is the reference's:

  * `RefSearch.playout_sequential()`  — MCTS.playout as the reference runs it (one playout at a time): Node fields
    (MCTS.py:12-19), expansion rules incl. the pass child and the single-move child with the literal prior 1
    (MCTS.py:108-121), select = first maximum of Q + u with U = c_puct * P * sqrt(N_parent) / (0.01 + n)
    (MCTS.py:39-49), leaf_value = (1 - lmbda) v + lmbda z (MCTS.py:123-125), same-sign running-mean backup to the
    root (MCTS.py:61-72).  Pinned against the reference's own trees in tests/golden/mcts.npz by replaying the
    logged v / z / priors (tests/test_oracle_golden.py).
  * `RefSearch.search(n, leaf_batch, virtual_loss)` — the batched algorithm of iago_b200/csrc/mcts.cu (waves of
    leaf_batch descents per tree, virtual visits, parked expansions, order-free fixed-point backup).  With
    leaf_batch = 1 it must build exactly the tree of playout_sequential(); with leaf_batch > 1 it is the checker
    for the GPU's batched mode (the GPU path is deterministic, so equality is exact).

The evaluators are injected: value_func(state, color) -> np.float32, rollout_func(state, color, playout_index) -> int,
policy_func(state, color) -> float32[64] probabilities.  Rules come from the C oracle (oracle/othello_ref.c).
"""
import numpy as np

from . import cref

FIX = float(1 << 40)


class RefNode:
    __slots__ = ("parent", "children", "n_visits", "Q", "P", "u", "W", "vn", "v", "pending", "claimed")

    def __init__(self, parent=None, prob=0):
        self.parent = parent
        self.children = {}      # action -> RefNode, insertion order = ascending action
        self.n_visits = 0
        self.Q = 0
        self.u = prob + 0.1     # MCTS.py:18
        self.P = prob + 0.1     # MCTS.py:19
        self.W = 0              # batched mode: sum of leaf values in 2^-40 fixed point
        self.vn = 0             # virtual visits
        self.v = None           # cached value-net output
        self.pending = None     # list of legal actions while the priors are being computed
        self.claimed = False


def _U(node, c_puct):
    return c_puct * node.P * np.sqrt(node.parent.n_visits) / (0.01 + node.n_visits)   # MCTS.py:48-49


class RefSearch:
    def __init__(self, state, color, value_func, rollout_func, policy_func, lmbda=0.5, c_puct=1, n_thr=15, cache_value=True):
        self.root = RefNode(None, 1.0)   # MCTS.py:81
        self.state = np.array(state, np.float32).reshape(8, 8).copy()
        self.color = int(color)
        self.value_func, self.rollout_func, self.policy_func = value_func, rollout_func, policy_func
        self.lmbda, self.c_puct, self.n_thr, self.cache_value = lmbda, c_puct, n_thr, cache_value
        self.done = 0

    # ------------------------------------------------------------------ the reference, one playout at a time
    def _expand(self, node, state, c):
        """MCTS.py:108-121. Returns the legal action list."""
        actions = cref.legal_actions(state, c)
        if len(actions) < 1:
            node.children[-1] = RefNode(node, 1)
        if len(actions) == 1:
            node.children[actions[0]] = RefNode(node, 1)
        elif len(actions) > 1:   # (with no legal move the reference also runs the policy and discards the result)
            prob = self.policy_func(state, c)
            for a in actions:
                node.children[a] = RefNode(node, prob[a])
        return actions

    def _leaf_value(self, state, color, k):
        v = self.value_func(state, color) if self.lmbda < 1 else 0
        z = self.rollout_func(state, color, k) if self.lmbda > 0 else 0
        return (1 - self.lmbda) * v + self.lmbda * z   # MCTS.py:123-125

    def playout_sequential(self):
        state, c, node = self.state.copy(), self.color, self.root
        while True:
            if not node.children:
                if node.n_visits >= self.n_thr:
                    self._expand(node, state, c)
                    continue
                leaf_value = self._leaf_value(state, c, self.done)
                nd = node
                while nd is not None:          # MCTS.py:68-72: same value, same sign, up to the root
                    nd.n_visits += 1
                    nd.Q += (leaf_value - nd.Q) / nd.n_visits
                    nd = nd.parent
                break
            for ch in node.children.values():
                ch.u = _U(ch, self.c_puct)
            action, node = max(node.children.items(), key=lambda an: an[1].Q + an[1].u)   # first maximum
            if action != -1:
                cref.place_stone(state, action, c)
            c = 3 - c
        self.done += 1

    # ------------------------------------------------------------------ the batched algorithm of csrc/mcts.cu
    def _score(self, ch, parent, exact, vloss):
        if exact:
            return ch.Q + _U(ch, self.c_puct)
        cp = self.c_puct * ch.P                       # float32 product when P is float32, float64 for the literal 1.1
        tot = ch.n_visits + ch.vn
        q = ((ch.W / FIX) - vloss * ch.vn) / tot if tot > 0 else 0.0
        n_parent = parent.n_visits + parent.vn - 1
        return q + float(cp) * np.sqrt(float(n_parent)) / (0.01 + tot)

    def _descend(self, node, state, c, exact, vloss, requests):
        """Walks down from `node`. Returns (status, node, state, c); status 'eval' or 'parked'."""
        while True:
            if not node.children:
                if node.pending is not None:
                    return "parked", node, state, c
                if node.n_visits >= self.n_thr:
                    actions = cref.legal_actions(state, c)
                    if len(actions) <= 1:
                        a = actions[0] if actions else -1
                        node.children[a] = RefNode(node, 1)
                        continue
                    node.pending = actions
                    requests.append((node, state.copy(), c))
                    return "parked", node, state, c
                return "eval", node, state, c
            best, best_a = None, None
            for a, ch in node.children.items():
                val = self._score(ch, node, exact, vloss)
                if best is None or val > best:
                    best, best_a = val, a
            node = node.children[best_a]
            node.vn += 1
            if best_a != -1:
                cref.place_stone(state, best_a, c)
            c = 3 - c

    def search(self, n_playouts, leaf_batch=1, virtual_loss=1.0):
        exact = leaf_batch == 1
        target = self.done + n_playouts
        while self.done < target:
            slots, requests = [], []
            nb = min(leaf_batch, target - self.done)
            for s in range(nb):
                self.root.vn += 1
                slots.append(list(self._descend(self.root, self.state.copy(), self.color, exact, virtual_loss, requests)))
            for node, st, c in requests:
                prob = self.policy_func(st, c)
                for a in node.pending:
                    node.children[a] = RefNode(node, prob[a])
                node.pending = None
            for sl in slots:
                if sl[0] == "parked":
                    sl[:] = self._descend(sl[1], sl[2], sl[3], exact, virtual_loss, requests)
                    assert sl[0] == "eval"
            # evaluation: one value per distinct leaf when cached, one rollout per slot
            vals = []
            for s, (_, node, st, c) in enumerate(slots):
                v = 0
                if self.lmbda < 1:
                    if self.cache_value:
                        if node.v is None:
                            node.v = self.value_func(st, c)
                        v = node.v
                    else:
                        v = self.value_func(st, c)
                z = self.rollout_func(st, c, self.done + s) if self.lmbda > 0 else 0
                vals.append((1 - self.lmbda) * v + self.lmbda * z)
            for (_, node, _, _), lv in zip(slots, vals):
                fix = int(np.rint(float(lv) * FIX))
                nd = node
                while nd is not None:
                    nd.n_visits += 1
                    if exact:
                        nd.Q += (lv - nd.Q) / nd.n_visits
                    nd.W += fix
                    nd.vn -= 1
                    nd = nd.parent
            self.done += nb

    # ------------------------------------------------------------------ results
    def best_move(self):
        return max(self.root.children.items(), key=lambda an: an[1].n_visits)[0]   # MCTS.py:147

    def update_with_move(self, last_move):
        """MCTS.py:149-154 (the caller keeps self.state / self.color in step)."""
        if last_move in self.root.children:
            self.root = self.root.children[last_move]
            self.root.parent = None
        else:
            self.root = RefNode(None, 1.0)

    def flatten(self, exact=True):
        """Breadth-first arrays, children in insertion (= ascending action) order — same form as the golden trees."""
        nodes, parent, action = [self.root], [-1], [0]
        i = 0
        while i < len(nodes):
            for a, ch in nodes[i].children.items():
                nodes.append(ch); parent.append(i); action.append(int(a))
            i += 1
        if exact:
            Q = np.array([float(nd.Q) for nd in nodes], np.float64)
        else:
            Q = np.array([nd.W / FIX / nd.n_visits if nd.n_visits else 0.0 for nd in nodes], np.float64)
        return dict(parent=np.array(parent, np.int32), action=np.array(action, np.int8),
                    n=np.array([nd.n_visits for nd in nodes], np.int32), Q=Q,
                    P=np.array([float(nd.P) for nd in nodes], np.float64))
