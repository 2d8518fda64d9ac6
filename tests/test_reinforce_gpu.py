"""K6 — the REINFORCE gradient and Adam step (iago_reinforce_*) vs the torch float64 autograd restatement of
src/train_rl.py:55-66 (oracle/reinforce_ref.py) on the reference's own recorded games (tests/golden/selfplay.npz).

Tolerance (fp32 arithmetic on the GPU, float64 in the oracle): per parameter tensor, max |g_gpu - g_ref| <= 1e-3 * max |g_ref|
(+1e-6 absolute); loss numerator relative 1e-5; the Adam + WeightDecay rule applied to the GPU's gradient within 1e-6 absolute per step."""
import numpy as np
import pytest

from conftest import load_golden, model_file

pytestmark = pytest.mark.gpu


def golden_batch():
    g = load_golden("selfplay")
    states, actions, rewards = [], [], []
    for i in range(len(g["seed"])):
        k = int(g["n_states"][i])
        states.append(g["states"][i][:k].reshape(k, 8, 8))
        actions.append(g["actions"][i][:k])
        rewards.append(np.full(k, g["judge"][i], np.float32))
    return np.concatenate(states), np.concatenate(actions).astype(np.int64), np.concatenate(rewards)


def to_device(states):
    import torch
    from iago_b200 import boards
    # recorded states are colour-swapped: learner's stones are 2, the opponent's 1 (rl_self_play.py:134-138)
    opp, own = boards.to_bitboards(states)
    t = lambda a: torch.from_numpy(a.view(np.int64).copy()).cuda()
    return t(own), t(opp)


def per_tensor_errors(flat_gpu, ref_grads):
    from oracle import reinforce_ref
    out, o = {}, 0
    for k in reinforce_ref.KEYS:
        n = ref_grads[k].size
        a, b = flat_gpu[o:o + n], ref_grads[k].reshape(-1)
        out[k] = (np.abs(a - b).max(), np.abs(b).max())
        o += n
    return out


def test_gradient_matches_autograd(engine):
    import torch
    from iago_b200 import npz
    from iago_b200.train_rl import ReinforceTrainer, N_PARAMS
    from oracle import nets, reinforce_ref
    path = model_file("RL/model2.npz")
    states, actions, rewards = golden_batch()
    tr = ReinforceTrainer(path, max_positions=256, tensor_cores=False)      # 242 positions -> one chunk; fp32 kernels
    own, opp = to_device(states)
    a = torch.from_numpy(actions.astype(np.int8)).cuda()
    r = torch.from_numpy(rewards).cuda()
    probs = tr.gradient(own, opp, a, r, want_probs=True)
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy().astype(np.float64)
    total, ref, pred = reinforce_ref.loss_and_grad(nets.load_params(path, np.float64), states, actions, rewards)
    assert g[N_PARAMS + 1] == len(states)
    assert abs(g[N_PARAMS] - total) <= 1e-5 * abs(total) + 1e-6
    assert np.abs(probs.cpu().numpy() - pred).max() <= 5e-5    # fp32 forward, logits up to ~130
    errs = per_tensor_errors(g[:N_PARAMS], ref)
    print("worst relative gradient error per tensor:", max(e / s_ for e, s_ in errs.values() if s_ > 0))
    for k, (e, scale) in errs.items():
        assert e <= 1e-3 * scale + 1e-6, (k, e, scale)
    # chunked accumulation (max_positions smaller than the batch) gives the same gradient
    tr2 = ReinforceTrainer(path, max_positions=100, slot=5, tensor_cores=False)
    tr2.gradient(own, opp, a, r)
    torch.cuda.synchronize()
    g2 = tr2.grad.cpu().numpy().astype(np.float64)
    assert g2[N_PARAMS + 1] == len(states)
    errs = per_tensor_errors(g2[:N_PARAMS], ref)
    for k, (e, scale) in errs.items():
        assert e <= 1e-3 * scale + 1e-6, (k, e, scale)
    # bit-reproducible run to run
    tr.gradient(own, opp, a, r)
    torch.cuda.synchronize()
    assert (tr.grad.cpu().numpy().astype(np.float64) == g).all()


def test_tensor_core_path(engine, mode=1):
    """tensor_cores=1: forward on the fused tcgen05 trunk (fp16 hi/lo split), data gradients as one fused tcgen05 chain (bf16
    hi/lo split), weight gradients as fp16 tcgen05 GEMMs.
    (1) The forward's activations equal the float64 forward within 2e-4; its ReLU on/off decisions differ from the float64
        ones ONLY on units whose pre-activation is within 2e-4 of zero (the kink, where the derivative is ambiguous).
    (2) With those on/off decisions taken as given, the gradient equals float64 autograd within 3e-3 * max|g| per tensor
        (measured 8e-4: fp16 weight-gradient operands with dY scaled and split hi/lo, fp32 accumulation); loss numerator within 1e-4 relative."""
    import torch
    from iago_b200 import npz
    from iago_b200.train_rl import ReinforceTrainer, N_PARAMS
    from oracle import nets, reinforce_ref
    path = model_file("RL/model2.npz")
    states, actions, rewards = golden_batch()
    own, opp = to_device(states)
    a = torch.from_numpy(actions.astype(np.int8)).cuda()
    r = torch.from_numpy(rewards).cuda()
    tr = ReinforceTrainer(path, max_positions=256, tensor_cores=mode, slot=4)
    tr.gradient(own, opp, a, r)
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy().astype(np.float64)
    col = torch.ones(own.numel(), dtype=torch.uint8, device="cuda")
    _, acts = engine.policy_forward_acts(tr.slot, own, opp, col)
    acts = [x.cpu().numpy() for x in acts]
    p64 = nets.load_params(path, np.float64)
    pre = []
    reinforce_ref.loss_and_grad(p64, states, actions, rewards, keep=pre)           # float64 pre-activations
    flips = 0
    for l in range(8):
        assert np.abs(acts[l] - np.maximum(pre[l], 0)).max() <= 2e-4
        diff = (acts[l] > 0) != (pre[l] > 0)
        flips += int(diff.sum())
        assert not diff.any() or np.abs(pre[l][diff]).max() <= 2e-4, l
    total, ref, _ = reinforce_ref.loss_and_grad(p64, states, actions, rewards, masks=[x > 0 for x in acts])
    assert abs(g[N_PARAMS] - total) <= 1e-4 * abs(total) + 1e-6
    errs = per_tensor_errors(g[:N_PARAMS], ref)
    worst = max(e / s_ for e, s_ in errs.values() if s_ > 0)
    print(f"tensor-core path (mode {mode}): {flips} ReLU decisions at the kink differ from float64; worst per-tensor gradient error {worst:.2e} of max|g|")
    for k, (e, scale) in errs.items():
        assert e <= 3e-3 * scale + 1e-6, (k, e, scale)
    tr.close()


def test_adam_steps_match_chainer_rule(engine):
    import torch
    from iago_b200 import npz
    from iago_b200.train_rl import ReinforceTrainer, N_PARAMS
    from oracle import nets, reinforce_ref
    path = model_file("RL/model2.npz")
    states, actions, rewards = golden_batch()
    tr = ReinforceTrainer(path, alpha=1e-3, max_positions=256, tensor_cores=False)
    own, opp = to_device(states)
    a = torch.from_numpy(actions.astype(np.int8)).cuda()
    r = torch.from_numpy(rewards).cuda()
    p64 = nets.load_params(path, np.float64)
    w = reinforce_ref.flat(p64)
    m, v, t = np.zeros_like(w), np.zeros_like(w), 0
    for step in range(2):
        tr.gradient(own, opp, a, r)
        torch.cuda.synchronize()
        g = tr.grad.cpu().numpy().astype(np.float64)      # the Adam rule is checked on the GPU's own gradient: at t = 1 the step is
        loss, count = tr.update()                         # alpha * sign(g), so gradient noise on near-zero entries would otherwise
        assert count == len(states)                       # dominate (gradient parity is the previous test)
        assert abs(loss - g[N_PARAMS] / count) <= 1e-6 * abs(loss) + 1e-9
        w, m, v, t = reinforce_ref.adam_step(w, g[:N_PARAMS] / count, m, v, t)
        got, gm, gv, gt = tr.state()
        assert gt == t
        assert np.abs(got - w).max() <= 1e-6, np.abs(got - w).max()
        assert np.abs(gm - m).max() <= 1e-6 * max(1.0, np.abs(m).max()) and np.abs(gv - v).max() <= 1e-6 * max(1.0, np.abs(v).max())
        w = got.astype(np.float64)                        # follow the fp32 replica, as a second rank would
    # the playing slot follows the update: the policy output of the trainer's slot equals the fp32 forward of the new weights
    from iago_b200 import boards
    p1, p2 = boards.to_bitboards(boards.start_state())
    out = engine.policy_forward_host(tr.slot, p1, p2, 1, probs=True)[0]
    new = npz.unflatten(w.astype(np.float32), npz.KIND_POLICY)
    ref_prob = nets.sl_policy({k: x.astype(np.float64) for k, x in new.items()}, nets.planes_from_state(boards.start_state()[None], 1, np.float64))[0]
    assert np.abs(out - ref_prob).max() <= 1e-4
    # ... and the device-side repack of the slot (iago_reinforce_sync_slot) is bit-identical to iago_load_net of the same weights
    engine.load_net(3, npz.unflatten(got, npz.KIND_POLICY))
    q1, q2 = boards.to_bitboards(states[:64])
    assert (engine.policy_forward_host(tr.slot, q1, q2, 1, probs=False) == engine.policy_forward_host(3, q1, q2, 1, probs=False)).all()


def test_train_set_runs_and_checkpoints(engine, tmp_path):
    from iago_b200 import network, npz
    from iago_b200.train_rl import ReinforceTrainer
    opp = network.SLPolicy().load(model_file("RL/model0.npz"))
    tr = ReinforceTrainer(model_file("RL/model2.npz"), max_positions=4096)
    before = tr.state()[0].copy()
    stats = tr.train_set(opp, n_games=64, seed=1)
    assert 0.0 <= stats["rate"] <= 1.0 and 64 * 20 <= stats["positions"] <= 64 * 34 and np.isfinite(stats["loss"])
    after, m, v, t = tr.state()
    assert t == 1 and np.abs(after - before).max() > 0 and np.abs(after - before).max() <= 1.1e-3   # |Adam step| <= alpha at t = 1
    tr.save_model(tmp_path / "model.npz")
    tr.save_optimizer(tmp_path / "opt.npz")
    z = np.load(tmp_path / "model.npz")
    assert sorted(z.files) == sorted(npz.TRUNK_KEYS + npz.HEAD_KEYS[npz.KIND_POLICY])
    o = np.load(tmp_path / "opt.npz")
    assert int(o["t"]) == 1 and o["block3/conv/W/m"].shape == (128, 128, 3, 3) and "bias10/b/v" in o.files
    again = network.SLPolicy().load(tmp_path / "model.npz")   # loadable like any reference archive
    assert again.params["conv9/W"].shape == (1, 128, 1, 1)


def test_gradient_is_bit_identical_from_run_to_run(engine):
    """Every reduction of the gradient runs in a fixed order (TMEM accumulation over a slice's positions, in-order sums over slices,
    taps and head slices), so repeating a gradient call must give the same bits — a missed hand-over between the producer warps, the
    MMA issuer and the CTA pairs of the backward chain would show up here (tools/stress_wgrad.py is the long form)."""
    import torch
    from iago_b200 import network
    from iago_b200.train_rl import ReinforceTrainer
    opp = network.SLPolicy().load(model_file("RL/model0.npz"))
    tr = ReinforceTrainer(model_file("RL/model2.npz"), max_positions=4096)
    d = tr.play_set(opp, 160, seed=11)
    total = d["own"].numel()
    assert total > 3000
    for n in (total, 2049, 333, 5):   # several chunkings: more positions than max_positions, ragged slices, fewer positions than slices
        ref = None
        for rep in range(4):
            tr.gradient(d["own"][:n], d["opp"][:n], d["action"][:n], d["reward"][:n])
            torch.cuda.synchronize()
            g = tr.grad.clone()
            if ref is None:
                ref = g
                assert torch.isfinite(g).all() and float(g.abs().max()) > 0
            else:
                assert torch.equal(g, ref), (n, rep, float((g - ref).abs().max()))


def test_checkpoint_loads_into_the_unmodified_reference_network(engine, tmp_path):
    """src/train_rl.py:73-79 writes snapshots with serializers.save_npz and reads them back with load_npz(path, model): a snapshot of
    this trainer (plain and 'predictor/'-prefixed, as models/rl_model.npz is) is loaded by the chainer stand-in's serializers.load_npz
    into the UNMODIFIED reference network.SLPolicy (baseline/_ref/network.py) in a separate interpreter, and that model's forward on
    real positions equals the GPU forward of the trainer's playing slot."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from iago_b200 import boards, network
    from iago_b200.train_rl import ReinforceTrainer
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref_dir, "network.py")):
        pytest.fail("baseline/_ref/network.py missing: run oracle/fetch_ref.py in the build container")
    opp = network.SLPolicy().load(model_file("RL/model0.npz"))
    tr = ReinforceTrainer(model_file("RL/model2.npz"), max_positions=4096, precision=3)
    tr.train_set(opp, n_games=64, seed=3)     # one real update, so the snapshot is not a file that already existed
    tr.save_model(tmp_path / "model.npz")
    tr.save_model(tmp_path / "model_pred.npz", prefix="predictor/")
    g = np.load(os.path.join(ROOT, "tests", "golden", "nets.npz"))
    x = g["x"][:96].astype(np.float32)
    np.save(tmp_path / "x.npy", x)
    code = f"""
import sys, json, numpy as np
sys.path[:0] = [{os.path.join(ROOT, 'oracle', 'chainer_shim')!r}, {ref_dir!r}]
import chainer
from chainer import serializers
import network                                   # the reference's file, unmodified
chainer.config.train = False
x = np.load({str(tmp_path / 'x.npy')!r})
out = {{}}
for name, path in (("plain", ""), ("pred", "predictor/")):
    m = network.SLPolicy()
    f = {str(tmp_path)!r} + ("/model.npz" if name == "plain" else "/model_pred.npz")
    if path:
        serializers.load_npz(f, m, path=path)
    else:
        serializers.load_npz(f, m)
    np.save({str(tmp_path)!r} + "/ref_" + name + ".npy", m(chainer.Variable(x)).data)
print("ok")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]
    own, opp_b = network.planes_to_bitboards(x)
    gpu = engine.policy_forward_host(tr.slot, own, opp_b, 1, probs=True, precision=3)
    for name in ("plain", "pred"):
        ref = np.load(tmp_path / f"ref_{name}.npy")
        assert ref.shape == gpu.shape and np.abs(ref - gpu).max() <= 1e-4, name
    # The forward half of the REINFORCE loss (src/train_rl.py:55-64) on the unmodified reference network: pred = model1(x) are
    # PROBABILITIES, c = softmax_cross_entropy(pred, y, reduce='no') applies log-softmax to them again, loss numerator = sum(c * r).
    # The trainer's own forward + loss (vector element [N_PARAMS] of iago_reinforce_grad) must give the same number.
    import torch
    from iago_b200.train_rl import N_PARAMS
    ref = np.load(tmp_path / "ref_plain.npy").astype(np.float64)
    rs = np.random.RandomState(5)
    y = rs.randint(0, 64, size=len(x))
    rew = rs.choice([-1.0, 0.0, 1.0], size=len(x))
    z = ref - ref.max(axis=1, keepdims=True)
    c = -(z[np.arange(len(x)), y] - np.log(np.exp(z).sum(axis=1)))          # Chainer's softmax_cross_entropy(reduce='no'), restated
    want = float((c * rew).sum())
    dev = tr.grad.device
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    probs = tr.gradient(t(own.view(np.int64), torch.int64), t(opp_b.view(np.int64), torch.int64), t(y, torch.int8), t(rew, torch.float32), want_probs=True)
    got = float(tr.grad[N_PARAMS])
    assert np.abs(probs.cpu().numpy() - ref).max() <= 1e-4
    assert abs(got - want) <= 1e-4 * max(1.0, abs(want)) and int(tr.grad[N_PARAMS + 1]) == len(x)
