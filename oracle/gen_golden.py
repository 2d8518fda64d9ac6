"""Generate tests/golden/*.npz by running the UNMODIFIED reference (under oracle/chainer_shim).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py [--out tests/golden] [--games 1000]

The reference has no tests and no golden vectors of its own, so these files are the
pins: every array below is the output of the reference's own functions.  Instrumentation is by
wrapping (logging) reference methods, never by changing what they compute.

Files:
  rules.npz     legal_actions / place_stone on reachable positions, pass positions and arbitrary boards;
                perft(1..6) from the reference rules
  simulate.npz  mcts_self_play.Simulate full games: per-game np.random seed, the uniforms that seed
                yields (RandomState(seed).random_sample), move list, final board, result
  nets.npz      SLPolicy (sl_model, rl_model), Value, RolloutPolicy outputs on harvested positions
  selfplay.npz  src/rl_self_play.Game trajectories (normal and 'head/tail switched' openings)
  selfgame.npz  self_play.SelfGame.get_position_self calls (unbound, on a stand-in object): probabilities, choice, draws consumed
  env.npz       rl_env.GameEnv.step sequences (numpy standing in for cupy): actions, opponent answers, uniforms consumed
  mcts.npz      MCTS.playout sequences: per-playout v / z / priors and the resulting tree
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402


def start_board():
    s = np.zeros([8, 8], dtype=np.float32)
    s[4, 3] = 1
    s[3, 4] = 1
    s[3, 3] = 2
    s[4, 4] = 2
    return s


def u8(x):
    return np.asarray(x).astype(np.uint8)


def legal_mask(actions):
    m = 0
    for a in actions:
        m |= 1 << int(a)
    return np.uint64(m)


def gen_simulate(mods, n_games, seed0, out):
    Sim = mods["mcts_self_play"].Simulate
    gf = mods["game"].GameFunctions
    orig_place, orig_legal = Sim.place_stone, Sim.legal_actions
    harvest = []  # (state u8[64], color, legal list)

    def place(self, state, action, color):
        self._moves.append(int(action))
        self._movers.append(int(color))
        return orig_place(self, state, action, color)

    def legal(self, color):
        acts = orig_legal(self, color)
        if self._harvest:
            harvest.append((u8(self.state).reshape(64).copy(), int(color), list(acts)))
        return acts

    Sim.place_stone, Sim.legal_actions = place, legal
    try:
        seeds, moves, movers, finals, results, nmoves, uniforms, starts, colors = [], [], [], [], [], [], [], [], []

        def run(state, color, seed, do_harvest):
            np.random.seed(seed)
            sim = Sim(state)
            sim._moves, sim._movers, sim._harvest = [], [], do_harvest
            r = sim(color)
            nxt = np.random.random_sample()
            u = np.random.RandomState(seed).random_sample(64)
            # exactly one uniform per stone placed, none per pass (mcts_self_play.py:106)
            assert nxt == u[len(sim._moves)], "np.random.choice consumed an unexpected number of draws"
            mv = np.full(64, -1, np.int8)
            mv[:len(sim._moves)] = sim._moves
            mr = np.zeros(64, np.int8)
            mr[:len(sim._movers)] = sim._movers
            seeds.append(seed); moves.append(mv); movers.append(mr); finals.append(u8(sim.state).reshape(64))
            results.append(r); nmoves.append(len(sim._moves)); uniforms.append(u)
            starts.append(u8(state).reshape(64)); colors.append(color)
            return sim

        # (i) from the opening, colour 1 first (BASELINE config 1)
        for g in range(n_games):
            run(start_board(), 1, seed0 + g, do_harvest=(g < 60))
        # (ii) from mid-game positions, both colours to move (what MCTS leaves look like)
        n_mid = max(8, n_games // 5)
        for g in range(n_mid):
            s = start_board()
            c = 1
            ply = 8 + (g * 7) % 40
            for a, who in zip(moves[g][:ply], movers[g][:ply]):
                if a < 0:
                    break
                gf.place_stone(s, int(a), int(who))
                c = 3 - int(who)
            if g % 3 == 2:
                c = 3 - c  # also the "wrong" side to move: exercises immediate passes
            run(s, c, seed0 + 100000 + g, do_harvest=(g < 30))
        out["simulate"] = dict(
            seed=np.array(seeds, np.int64), start=np.array(starts, np.uint8), color=np.array(colors, np.int8),
            uniforms=np.array(uniforms, np.float64), moves=np.array(moves, np.int8),
            movers=np.array(movers, np.int8), n_moves=np.array(nmoves, np.int32),
            final=np.array(finals, np.uint8), result=np.array(results, np.int8))
    finally:
        Sim.place_stone, Sim.legal_actions = orig_place, orig_legal
    return harvest


def gen_rules(mods, harvest, out, rng):
    gf = mods["game"].GameFunctions
    states, colors, masks = [], [], []
    ps_state, ps_color, ps_action, ps_after = [], [], [], []

    def add(state64, color, n_place, any_cell=0):
        s = state64.reshape(8, 8).astype(np.float32)
        acts = gf.legal_actions(s, color)
        states.append(u8(state64)); colors.append(color); masks.append(legal_mask(acts))
        todo = list(acts) if n_place is None else list(rng.permutation(acts)[:n_place])
        todo += [int(a) for a in rng.integers(0, 64, size=any_cell)]  # place_stone has no legality check
        for a in todo:
            t = s.copy()
            r = gf.place_stone(t, int(a), color)
            assert r is t
            ps_state.append(u8(state64)); ps_color.append(color); ps_action.append(int(a)); ps_after.append(u8(t).reshape(64))

    seen = set()
    for st, c, _ in harvest:
        key = (st.tobytes(), c)
        if key in seen:
            continue
        seen.add(key)
        add(st, c, None)
        add(st, 3 - c, 2)
    # arbitrary (mostly unreachable) boards at every fill density
    for i in range(1500):
        fill = rng.random()
        r = rng.random(64)
        st = np.where(r < fill * 0.5, 1, np.where(r < fill, 2, 0)).astype(np.uint8)
        if i % 50 == 0:
            st[:] = [0, 1, 2][(i // 50) % 3]
        add(st, 1 + i % 2, 3, any_cell=2)
    # action -1 is a no-op (game.py:181-182)
    s = start_board()
    assert gf.place_stone(s, -1, 1) is s and (s == start_board()).all()

    def perft(s, color, depth, passed=False):
        if depth == 0:
            return 1
        acts = gf.legal_actions(s, color)
        if not acts:
            return 1 if passed else perft(s, 3 - color, depth - 1, True)
        tot = 0
        for a in acts:
            t = s.copy()
            gf.place_stone(t, a, color)
            tot += perft(t, 3 - color, depth - 1)
        return tot

    out["rules"] = dict(
        state=np.array(states, np.uint8), color=np.array(colors, np.int8), legal_mask=np.array(masks, np.uint64),
        ps_state=np.array(ps_state, np.uint8), ps_color=np.array(ps_color, np.int8),
        ps_action=np.array(ps_action, np.int8), ps_after=np.array(ps_after, np.uint8),
        perft=np.array([perft(start_board(), 1, d) for d in range(1, 7)], np.int64),
        start_legal_1=np.array(gf.legal_actions(start_board(), 1), np.int8),
        start_legal_2=np.array(gf.legal_actions(start_board(), 2), np.int8))


def gen_nets(mods, harvest, out, rng, n_pos=768):
    net, ser, gf = mods["network"], mods["chainer"].serializers, mods["game"].GameFunctions
    idx = rng.permutation(len(harvest))[:n_pos]
    st = np.array([harvest[i][0] for i in idx], np.uint8)
    col = np.array([harvest[i][1] for i in idx], np.int8)
    st = np.concatenate([start_board().astype(np.uint8).reshape(1, 64)] * 2 + [st])
    col = np.concatenate([np.array([1, 2], np.int8), col])
    x = np.concatenate([gf.make_state_var(s.reshape(8, 8).astype(np.float32), int(c)).data for s, c in zip(st, col)])
    sl = net.SLPolicy(); ser.load_npz("./models/sl_model.npz", sl)
    rl = net.SLPolicy(); ser.load_npz("./models/rl_model.npz", rl, path="predictor/")
    va = net.Value(); ser.load_npz("./models/value_model.npz", va)
    ro = net.RolloutPolicy(); ser.load_npz("./models/rollout_model.npz", ro)
    B = 64
    cat = lambda f: np.concatenate([f(x[i:i + B]).data for i in range(0, len(x), B)])
    legal = np.array([legal_mask(gf.legal_actions(s.reshape(8, 8).astype(np.float32), int(c))) for s, c in zip(st, col)])
    out["nets"] = dict(state=st, color=col, x=x.astype(np.uint8), legal_mask=legal,
                       sl_prob=cat(sl), rl_prob=cat(rl), value=cat(va), rollout_prob=cat(ro))


def gen_selfplay(mods, out, seed0):
    import random
    net, ser = mods["network"], mods["chainer"].serializers
    Game = mods["rl_self_play"].Game
    m1 = net.SLPolicy(); ser.load_npz("./models/RL/model2.npz", m1)
    m2 = net.SLPolicy(); ser.load_npz("./models/RL/model0.npz", m2)
    orig_place = Game.place_stone

    def place(self, state, action, color):
        self._moves.append(int(action)); self._movers.append(int(color))
        return orig_place(self, state, action, color)

    Game.place_stone = place
    try:
        rec = dict(seed=[], extra=[], moves=[], movers=[], n_moves=[], uniforms=[], final=[], judge=[],
                   n_states=[], states=[], actions=[])
        for g in range(8):
            seed = seed0 + g
            np.random.seed(seed); random.seed(seed)
            game = Game(m1, m2)
            game._moves, game._movers = [], []
            extra = -1
            if g % 2 == 1:  # src/train_rl.py:43-46 'switch head and tail' (no flip, no stone_num += 1)
                pos = random.choice([[2, 4], [3, 5], [4, 2], [5, 3]])
                game.state[pos[0], pos[1]] = 2
                extra = pos[0] * 8 + pos[1]
            states, actions, judge = game()
            nxt = np.random.random_sample()
            u = np.random.RandomState(seed).random_sample(64)
            assert nxt == u[len(game._moves)]
            mv = np.full(64, -1, np.int8); mv[:len(game._moves)] = game._moves
            mr = np.zeros(64, np.int8); mr[:len(game._movers)] = game._movers
            ss = np.zeros((32, 64), np.uint8); ss[:len(states)] = u8(np.array(states)).reshape(-1, 64)
            aa = np.full(32, -1, np.int8); aa[:len(actions)] = actions
            for k, v in dict(seed=seed, extra=extra, moves=mv, movers=mr, n_moves=len(game._moves), uniforms=u,
                             final=u8(game.state).reshape(64), judge=judge, n_states=len(states), states=ss,
                             actions=aa).items():
                rec[k].append(v)
        out["selfplay"] = {k: np.array(v) for k, v in rec.items()}
    finally:
        Game.place_stone = orig_place


def gen_env(mods, out, seed0):
    """rl_env.GameEnv (UNMODIFIED, imported with numpy standing in for cupy): learner plays scripted legal actions, the
    opponent (RL/model0 wrapped like L.Classifier) answers; logs per-step opponent probabilities and the uniforms consumed."""
    import importlib
    import types
    net, ser = mods["network"], mods["chainer"].serializers
    rl_env = importlib.import_module("rl_env")
    gf = mods["game"].GameFunctions
    m2 = net.SLPolicy(); ser.load_npz("./models/RL/model0.npz", m2)
    wrapped = types.SimpleNamespace(predictor=m2)
    rec = dict(seed=[], actions=[], opp_actions=[], n_steps=[], uniforms=[], n_draws=[], final=[], judge=[], done_step=[], probs=[])
    for g in range(6):
        seed = seed0 + g
        np.random.seed(seed)
        env = rl_env.GameEnv(None, wrapped)
        env.reset()
        lrng = np.random.default_rng(seed)      # the learner's scripted choices (not part of the reference's streams)
        acts, opps, probs = [], [], []
        done, steps = False, 0
        while not done and steps < 40:
            legal = gf.legal_actions(env.state.copy(), 1)
            a = int(legal[lrng.integers(len(legal))]) if legal else 0
            before = env.state.copy()
            obs, r, done, info = env.step(a)
            # opponent's move = the colour-2 stone that appeared on a previously empty cell
            mid = before.copy()
            if legal:
                gf.place_stone(mid, a, 1)
            new2 = np.argwhere((env.state == 2) & (mid == 0))
            opps.append(int(new2[0][0] * 8 + new2[0][1]) if len(new2) else -1)
            pr = m2(np.stack([mid == 1, mid == 2], axis=0).astype(np.float32).reshape(1, 2, 8, 8)).data.reshape(64)
            probs.append(pr.astype(np.float32))
            acts.append(a)
            steps += 1
        nxt = np.random.random_sample()
        u = np.random.RandomState(seed).random_sample(4096)
        nd = int(np.argmax(u == nxt))
        assert u[nd] == nxt
        pad = lambda v, fill: np.concatenate([np.array(v, np.int8), np.full(40 - len(v), fill, np.int8)])
        pp = np.zeros((40, 64), np.float32); pp[:len(probs)] = np.array(probs)
        for k, v in dict(seed=seed, actions=pad(acts, -1), opp_actions=pad(opps, -1), n_steps=steps, uniforms=u[:512], n_draws=nd,
                         final=u8(env.state).reshape(64), judge=env(), done_step=steps, probs=pp).items():
            rec[k].append(v)
    out["env"] = {k: np.array(v) for k, v in rec.items()}


def gen_selfgame(mods, harvest, out, seed0, n_calls=240):
    """self_play.SelfGame.get_position_self (UNMODIFIED, self_play.py:8-30).  The class cannot be instantiated at the reference HEAD
    (SelfGame() omits Game.__init__'s required argument and Game lost valid_pos / place_stone(position, color)), so the function is
    called unbound on a stand-in object that carries exactly what it reads: `state`, `model1.predictor`, `model2.predictor`.
    Recorded per call: the board before the call, colour, the legal list handed in, the model's probabilities for the encoded
    input, the position returned and how many np.random uniforms the call consumed (one per attempt of its rejection loop)."""
    import importlib
    import types
    net, ser, chainer = mods["network"], mods["chainer"].serializers, mods["chainer"]
    sp = importlib.import_module("self_play")
    gf = mods["game"].GameFunctions
    m1 = net.SLPolicy(); ser.load_npz("./models/sl_model.npz", m1)
    m2 = net.SLPolicy(); ser.load_npz("./models/RL/model0.npz", m2)
    rec = dict(seed=[], state=[], color=[], legal_mask=[], probs=[], action=[], n_draws=[], uniforms=[])
    picks = [h for h in harvest if len(h[2]) >= 1][::max(1, len(harvest) // n_calls)][:n_calls]
    skipped = 0
    for k, (st, c, legal) in enumerate(picks):
        seed = seed0 + k
        stub = types.SimpleNamespace(state=st.reshape(8, 8).astype(np.float32).copy(), model1=types.SimpleNamespace(predictor=m1),
                                     model2=types.SimpleNamespace(predictor=m2))
        stub.get_position_self = types.MethodType(sp.SelfGame.get_position_self, stub)   # the function recurses through self
        positions = [[a // 8 + 1, a % 8 + 1] for a in legal]
        # the model's view of the position (what the function feeds its predictor): colour 1 sees the swapped board
        view = stub.state.copy()
        if c == 1:
            view = view * (3 - view) * (3 - view) / 2
        X = np.stack([view == 1, view == 2], axis=0).astype(np.float32).reshape(2, 1, 8, 8).transpose(1, 0, 2, 3)
        pr = (m1 if c == 1 else m2)(chainer.Variable(X)).data.reshape(64).astype(np.float32).copy()
        np.random.seed(seed)
        try:
            pos = stub.get_position_self(c, positions)
        except RecursionError:      # the reference's own failure mode when the net's mass sits on illegal cells: not recorded
            skipped += 1
            continue
        nxt = np.random.random_sample()
        u = np.random.RandomState(seed).random_sample(4096)
        nd = int(np.argmax(u == nxt))
        assert u[nd] == nxt and pos in positions
        for key, v in dict(seed=seed, state=u8(st).reshape(64), color=c, legal_mask=legal_mask(legal), probs=pr,
                           action=(pos[0] - 1) * 8 + pos[1] - 1, n_draws=nd, uniforms=u[:64]).items():
            rec[key].append(v)
        assert nd <= 64
    print("selfgame:", len(rec["seed"]), "calls recorded,", skipped, "ended in the reference's RecursionError", flush=True)
    out["selfgame"] = {k: np.array(v) for k, v in rec.items()}


def gen_valuegen(mods, out, seed0, n_games=24):
    """value_self_play.SelfPlay (UNMODIFIED).  The file is dead at the reference HEAD: it imports a module `SLPolicy` that no
    longer exists and loads un-prefixed archives into L.Classifier.  Stand-ins, none of which touch the game logic:
      * module SLPolicy with SLPolicyNet = network.SLPolicy minus its final softmax (the file applies its own softmax to the
        net output, value_self_play.py:142) — F.softmax is the identity while that net runs;
      * serializers.load_npz into the Classifier's predictor (rl_model.npz carries the 'predictor/' prefix, sl_model.npz not);
      * chainer.functions.loss.softmax_cross_entropy importable;
      * random.choice(seq) -> seq[floor(u * len)] with u the NEXT np.random uniform, so one recorded stream drives the game.
    Games in which the reference's un-stabilised softmax overflows (np.random.choice raises on NaN) are skipped."""
    import importlib
    import random
    import sys
    import types
    net, chainer = mods["network"], mods["chainer"]

    log = []

    class SLPolicyNet(net.SLPolicy):
        def __call__(self, x):
            keep = net.F.softmax
            net.F.softmax = lambda h, axis=1: h
            try:
                y = super().__call__(x)
            finally:
                net.F.softmax = keep
            log.append(np.array(y.data, np.float32).reshape(64))
            return y

    sys.modules["SLPolicy"] = types.SimpleNamespace(SLPolicyNet=SLPolicyNet)
    loss = types.ModuleType("chainer.functions.loss")
    sce = types.ModuleType("chainer.functions.loss.softmax_cross_entropy")
    sce.softmax_cross_entropy = lambda *a, **k: None
    loss.softmax_cross_entropy = sce
    sys.modules["chainer.functions.loss"] = loss
    sys.modules["chainer.functions.loss.softmax_cross_entropy"] = sce
    ser = chainer.serializers
    real_load = ser.load_npz

    def load_npz(path, obj, *a, **k):
        target = getattr(obj, "predictor", obj)
        with np.load(path) as z:
            pre = "predictor/" if any(f.startswith("predictor/") for f in z.files) else ""
        return real_load(path, target, pre) if pre else real_load(path, target)

    vsp = importlib.import_module("value_self_play")
    vsp.serializers.load_npz = load_npz
    real_choice = random.choice
    random.choice = lambda seq: seq[min(int(np.random.random_sample() * len(seq)), len(seq) - 1)]
    rec = dict(seed=[], stop_num=[], state=[], result=[], uniforms=[], n_draws=[], final=[], logits=[], n_logits=[])
    skipped = 0
    try:
        g = 0
        while len(rec["seed"]) < n_games:
            seed = seed0 + g
            g += 1
            stop = 4 + (seed * 7) % 60
            np.random.seed(seed)
            del log[:]
            sp = vsp.SelfPlay(stop)
            try:
                state, result = sp()
            except ValueError:
                skipped += 1
                continue
            nxt = np.random.random_sample()
            u = np.random.RandomState(seed).random_sample(512)
            nd = int(np.argmax(u == nxt))
            assert u[nd] == nxt
            lg = np.zeros((64, 64), np.float32)
            lg[:len(log)] = np.array(log)
            for k, v in dict(seed=seed, stop_num=stop, state=u8(state).reshape(64), result=result, uniforms=u[:160], n_draws=nd,
                             final=u8(sp.state).reshape(64), logits=lg,
                             n_logits=len(log)).items():
                rec[k].append(v)
    finally:
        random.choice = real_choice
        vsp.serializers.load_npz = real_load
    print("valuegen: skipped", skipped, "games whose softmax overflowed in the reference")
    out["valuegen"] = {k: np.array(v) for k, v in rec.items()}


def gen_load(out, seed=31337, n_lines=140):
    """load.py (UNMODIFIED) main() on a synthetic policy_data/txt/data.txt: the shuffled npy files it writes."""
    import importlib.util
    import tempfile
    rs = np.random.RandomState(seed)
    lines = []
    for i in range(n_lines):
        cells = rs.randint(0, 3, size=64)
        lines.append(" ".join(str(int(c)) for c in cells) + f" {rs.randint(1, 9)} {rs.randint(1, 9)} {'BW'[rs.randint(2)]}\n")
    root = tempfile.mkdtemp()
    os.makedirs(os.path.join(root, "policy_data", "txt")); os.makedirs(os.path.join(root, "policy_data", "npy")); os.makedirs(os.path.join(root, "run"))
    with open(os.path.join(root, "policy_data", "txt", "data.txt"), "w") as f:
        f.writelines(lines)
    spec = importlib.util.spec_from_file_location("ref_load", os.path.join(ref_harness.REF, "load.py"))
    ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
    cwd = os.getcwd()
    os.chdir(os.path.join(root, "run"))
    try:
        np.random.seed(seed)
        ref.main()
    finally:
        os.chdir(cwd)
    ld = lambda n: np.load(os.path.join(root, "policy_data", "npy", n))
    out["load"] = dict(seed=np.array(seed), lines=np.array(lines), states=ld("states.npy").astype(np.int8), actions=ld("actions.npy"),
                       states_test=ld("states_test.npy").astype(np.int8), actions_test=ld("actions_test.npy"),
                       rotate=ref.rotate(np.arange(64.0)), transpose=ref.transpose(np.arange(64.0)))


def flatten_tree(root):
    """BFS; children in dict insertion order (= ascending action, as expand() inserts them)."""
    nodes, parent, action = [root], [-1], [0]
    i = 0
    while i < len(nodes):
        for a, ch in nodes[i].children.items():
            nodes.append(ch); parent.append(i); action.append(int(a))
        i += 1
    return dict(parent=np.array(parent, np.int32), action=np.array(action, np.int8),
                n=np.array([nd.n_visits for nd in nodes], np.int32),
                Q=np.array([float(nd.Q) for nd in nodes], np.float64),
                P=np.array([float(nd.P) for nd in nodes], np.float64),
                q_is_f32=np.array([isinstance(nd.Q, np.float32) for nd in nodes], np.bool_))


def gen_mcts(mods, out, sim_rec):
    M, gf = mods["MCTS"], mods["game"].GameFunctions
    cases = []

    def position(game, ply):
        s, c = start_board(), 1
        for a, who in zip(sim_rec["moves"][game][:ply], sim_rec["movers"][game][:ply]):
            gf.place_stone(s, int(a), int(who)); c = 3 - int(who)
        return s, c

    s19 = start_board(); gf.place_stone(s19, 19, 1)
    specs = [("after19", s19, 2, dict(), 400),
             ("mid30", *position(3, 30), dict(), 300),
             ("late52_lam1_thr2", *position(5, 52), dict(lmbda=1.0, n_thr=2), 300),
             ("late56_lam0_thr1", *position(7, 56), dict(lmbda=0.0, n_thr=1), 200)]
    for name, state, color, kw, n_play in specs:
        m = M.MCTS(**kw)
        log = dict(v=[], z=[], prior_state=[], prior_color=[], prior=[])
        vf, rf, pf = m.value_func, m.evaluate_rollout, m.policy_func
        cur = {}

        def value_func(st, c, vf=vf, cur=cur):
            cur["v"] = vf(st, c); return cur["v"]

        def evaluate_rollout(st, c, rf=rf, cur=cur):
            cur["z"] = rf(st, c); return cur["z"]

        def policy_func(st, c, actions, pf=pf, log=log):
            ap = pf(st, c, actions)
            pr = np.zeros(64, np.float32)
            for a, p in ap:
                pr[a] = p
            log["prior_state"].append(u8(st).reshape(64).copy()); log["prior_color"].append(c); log["prior"].append(pr)
            return ap

        m.value_func, m.evaluate_rollout, m.policy_func = value_func, evaluate_rollout, policy_func
        for k in range(n_play):
            np.random.seed(7000 + k)
            cur.clear()
            m.playout(state.copy(), color, m.root)
            log["v"].append(float(cur.get("v", 0.0))); log["z"].append(int(cur.get("z", 0)))
        tree = flatten_tree(m.root)
        best = max(m.root.children.items(), key=lambda an: an[1].n_visits)[0]
        case = dict(root_state=u8(state).reshape(64), root_color=np.int8(color),
                    lmbda=np.float64(m.lmbda), c_puct=np.float64(m.c_puct), n_thr=np.int32(m.n_thr),
                    n_playouts=np.int32(n_play), v=np.array(log["v"], np.float32), z=np.array(log["z"], np.int8),
                    prior_state=np.array(log["prior_state"], np.uint8).reshape(-1, 64),
                    prior_color=np.array(log["prior_color"], np.int8),
                    prior=np.array(log["prior"], np.float32).reshape(-1, 64), best=np.int8(best),
                    **{"tree_" + k: v for k, v in tree.items()})
        cases.append((name, case))
    flat = {"cases": np.array([n for n, _ in cases])}
    for n, c in cases:
        for k, v in c.items():
            flat[f"{n}/{k}"] = v
    out["mcts"] = flat


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden"))
    ap.add_argument("--games", type=int, default=1000)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    outdir = os.path.abspath(args.out)
    mods = ref_harness.load()
    rng = np.random.default_rng(20261017)
    out = {}
    harvest = gen_simulate(mods, args.games, 12345, out)
    print("simulate done", len(harvest), "harvested positions", flush=True)
    only = set(args.only.split(",")) if args.only else None
    if not only or "rules" in only:
        gen_rules(mods, harvest, out, rng); print("rules done", flush=True)
    if not only or "nets" in only:
        gen_nets(mods, harvest, out, rng); print("nets done", flush=True)
    if not only or "selfplay" in only:
        gen_selfplay(mods, out, 777); print("selfplay done", flush=True)
    if not only or "env" in only:
        gen_env(mods, out, 4242); print("env done", flush=True)
    if not only or "selfgame" in only:
        gen_selfgame(mods, harvest, out, 31000); print("selfgame done", flush=True)
    if not only or "valuegen" in only:
        gen_valuegen(mods, out, 9090); print("valuegen done", flush=True)
    if not only or "load" in only:
        gen_load(out); print("load done", flush=True)
    if not only or "mcts" in only:
        gen_mcts(mods, out, out["simulate"]); print("mcts done", flush=True)
    os.makedirs(outdir, exist_ok=True)
    for name, d in out.items():
        if only and name not in only and not (name == "simulate" and "simulate" in only):
            continue
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **d)
        print("wrote", name, {k: getattr(v, "shape", None) for k, v in list(d.items())[:6]})


if __name__ == "__main__":
    main()
