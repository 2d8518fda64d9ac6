"""Supervised trainers (train_policy.py / train_value.py of the package) vs torch float64 autograd (oracle/supervised_ref.py,
oracle/reinforce_ref.py).  Tolerances as for K6 (tests/test_reinforce_gpu.py): fp32 path 1e-3 * max|g| per tensor; tensor-core
path 3e-3 * max|g| with the forward's own ReLU decisions taken as given."""
import numpy as np
import pytest

from conftest import load_golden, model_file

pytestmark = pytest.mark.gpu


def records(n=192, seed=0):
    """Positions of the reference's own self-play games (tests/golden/selfplay.npz) as supervised records."""
    g = load_golden("selfplay")
    states = np.concatenate([g["states"][i][:int(g["n_states"][i])].reshape(-1, 8, 8) for i in range(len(g["seed"]))])
    actions = np.concatenate([g["actions"][i][:int(g["n_states"][i])] for i in range(len(g["seed"]))]).astype(np.int64)
    rs = np.random.RandomState(seed)
    idx = rs.permutation(len(states))[:n]
    return states[idx].astype(np.float32), actions[idx], rs.choice([-1.0, 0.0, 1.0], size=len(idx)).astype(np.float32)


def test_rollout_policy_gradient_and_adam(engine, rollout_weights):
    import torch
    from iago_b200.train_policy import RolloutTrainer, states_to_device
    from oracle import reinforce_ref, supervised_ref
    W, b = rollout_weights
    states, actions, _ = records()
    tr = RolloutTrainer(W, b)
    own, opp = states_to_device(states, torch.device("cuda", 0))
    act = torch.from_numpy(actions.astype(np.int8)).cuda()
    tr.gradient(own, opp, act)
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy().astype(np.float64)
    total, dW, db, pred = supervised_ref.rollout_loss_and_grad(W, b, states, actions)
    assert g[83] == len(states) and abs(g[82] - total) <= 1e-5 * abs(total)
    assert np.abs(g[:18] - dW.reshape(18)).max() <= 1e-4 * np.abs(dW).max()
    assert np.abs(g[18:82] - db).max() <= 1e-4 * np.abs(db).max()
    # test-set metrics: mean double-softmax CE and accuracy (train_policy.py:69-70)
    loss, acc = tr.evaluate(own, opp, act)
    assert abs(loss - total / len(states)) <= 1e-5 * abs(loss) and acc == (pred.argmax(axis=1) == actions).mean()
    # Adam + WeightDecay on the GPU's gradient
    w0 = np.concatenate([W.reshape(18), b]).astype(np.float64)
    loss_mean, count = tr.update()
    w1, m1, v1, t1 = reinforce_ref.adam_step(w0, g[:82] / count, np.zeros(82), np.zeros(82), 0)
    got, gm, gv, gt = tr.state()
    assert gt == 1 and np.abs(got - w1).max() <= 1e-6 and np.abs(gm - m1).max() <= 1e-6 * max(1, np.abs(m1).max())
    # checkpoint in the reference's archive layout (models/rollout_model.npz / rollout_optimizer.npz)
    ref_opt_keys = ["bias2/b/m", "bias2/b/t", "bias2/b/v", "conv1/W/m", "conv1/W/t", "conv1/W/v", "epoch", "t"]   # models/rollout_optimizer.npz
    import tempfile, os
    d = tempfile.mkdtemp()
    tr.save_model(os.path.join(d, "m.npz")); tr.save_optimizer(os.path.join(d, "o.npz"))
    assert sorted(np.load(os.path.join(d, "m.npz")).files) == ["bias2/b", "conv1/W"]
    o = np.load(os.path.join(d, "o.npz"))
    assert sorted(o.files) == ref_opt_keys and o["conv1/W/m"].shape == (1, 2, 3, 3) and o["bias2/b/v"].shape == (64,)
    tr.close()


def test_sl_policy_step_is_cross_entropy(engine):
    """reward 1 for every record turns the K6 gradient into train_policy.py's loss; fp32 path against autograd."""
    import torch
    from iago_b200.train_policy import SLTrainer, states_to_device
    from iago_b200.train_rl import N_PARAMS
    from oracle import nets, reinforce_ref
    path = model_file("RL/model2.npz")
    states, actions, _ = records(128)
    tr = SLTrainer(path, max_positions=128, tensor_cores=False, slot=5)
    own, opp = states_to_device(states, torch.device("cuda", 0))
    act = torch.from_numpy(actions.astype(np.int8)).cuda()
    tr.gradient(own, opp, act, torch.ones(len(states), dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy().astype(np.float64)
    total, ref, pred = reinforce_ref.loss_and_grad(nets.load_params(path, np.float64), states, actions, np.ones(len(states)))
    assert abs(g[N_PARAMS] - total) <= 1e-5 * abs(total)
    o = 0
    for k in reinforce_ref.KEYS:
        n = ref[k].size
        assert np.abs(g[o:o + n] - ref[k].reshape(-1)).max() <= 1e-3 * np.abs(ref[k]).max() + 1e-6, k
        o += n
    loss, acc = tr.evaluate(own, opp, act)
    assert abs(loss - total / len(states)) <= 1e-4 * abs(loss) and abs(acc - (pred.argmax(axis=1) == actions).mean()) < 1e-9
    tr.close()


@pytest.mark.parametrize("tc", [False, True])
def test_value_gradient(engine, tc):
    import torch
    from iago_b200 import npz
    from iago_b200.train_policy import states_to_device
    from iago_b200.train_value import N_PARAMS, ValueTrainer
    from oracle import nets, supervised_ref
    path = model_file("value_model.npz")
    states, _, targets = records(160, seed=3)
    tr = ValueTrainer(path, max_positions=256, tensor_cores=tc, slot=7, seed=11)
    own, opp = states_to_device(states, torch.device("cuda", 0))
    y = torch.from_numpy(targets).cuda()
    pred, mask = tr.gradient(own, opp, y, want_pred=True, want_mask=True, position_id0=1000)
    torch.cuda.synchronize()
    g = tr.grad.cpu().numpy().astype(np.float64)
    mask = mask.cpu().numpy()
    assert 0.5 < mask.mean() < 0.7                                    # keep probability 0.6
    p64 = nets.load_params(path, np.float64)
    relu_masks = None
    if tc:
        col = torch.ones(own.numel(), dtype=torch.uint8, device="cuda")
        # the forward's own ReLU decisions (its activations differ from float64 by ~1e-4 at the kink)
        _, acts = engine.value_forward_acts(tr.slot, own, opp, col)
        torch.cuda.synchronize()
        relu_masks = [a.cpu().numpy() > 0 for a in acts]
    total, ref, v = supervised_ref.value_loss_and_grad(p64, states, targets, drop_mask=mask, relu_masks=relu_masks)
    tol = 3e-3 if tc else 1e-3
    assert np.abs(pred.cpu().numpy() - v).max() <= (2e-3 if tc else 1e-4) * max(1.0, np.abs(v).max())
    assert g[N_PARAMS + 1] == len(states) and abs(g[N_PARAMS] - total) <= (1e-3 if tc else 1e-5) * abs(total)
    o, worst = 0, 0.0
    for k in supervised_ref.VALUE_KEYS:
        n = ref[k].size
        e, scale = np.abs(g[o:o + n] - ref[k].reshape(-1)).max(), np.abs(ref[k]).max()
        worst = max(worst, e / scale if scale > 0 else 0.0)
        assert e <= tol * scale + 1e-6, (k, e, scale)
        o += n
    print(f"value gradient (tensor cores {tc}): worst per-tensor error {worst:.2e} of max|g|")
    # evaluation = dropout off; equals the float64 forward
    mse = tr.evaluate(own, opp, y)
    total0, _, v0 = supervised_ref.value_loss_and_grad(p64, states, targets, drop_mask=None)
    assert abs(mse - total0 / len(states)) <= 1e-3 * abs(total0 / len(states)) + 1e-6
    # one Adam step keeps the playing slot in sync with the new parameters (device-side repack of the value slot)
    tr.update()
    after = tr.params()
    engine.load_net(3, after)
    q1, q2 = own[:32].cpu().numpy().view(np.uint64), opp[:32].cpu().numpy().view(np.uint64)
    assert (engine.value_forward_host(tr.slot, q1, q2, 1) == engine.value_forward_host(3, q1, q2, 1)).all()
    tr.close()


def test_training_loops_run(engine, rollout_weights, tmp_path):
    """Two epochs of each loop on a small synthetic set: losses finite, rollout loss decreases, archives written."""
    from iago_b200 import train_policy, train_value
    states, actions, targets = records(192, seed=5)
    np.random.seed(0)
    tr, hist = train_policy.train(states[:160], actions[:160], states[160:], actions[160:], policy="rollout", epochs=3, minibatch=64,
                                  model_path=str(tmp_path / "r.npz"), optimizer_path=str(tmp_path / "ro.npz"), log=str(tmp_path / "r.txt"))
    assert all(np.isfinite(h[0]) for h in hist) and hist[-1][0] <= hist[0][0] + 1e-6 and (tmp_path / "r.npz").exists()
    tr.close()
    tr, hist = train_policy.train(states[:160], actions[:160], states[160:], actions[160:], policy="sl", epochs=1, minibatch=64,
                                  init=model_file("RL/model2.npz"), model_path=str(tmp_path / "s.npz"), log=str(tmp_path / "s.txt"))
    assert np.isfinite(hist[0][0]) and 0.0 <= hist[0][1] <= 1.0
    tr.close()
    tr, hist = train_value.train(states[:160], targets[:160], states[160:], targets[160:], epochs=1, minibatch=64,
                                 init=model_file("value_model.npz"), model_path=str(tmp_path / "v.npz"), optimizer_path=str(tmp_path / "vo.npz"),
                                 log=str(tmp_path / "v.txt"))
    assert np.isfinite(hist[0]) and sorted(np.load(tmp_path / "v.npz").files) == sorted(train_value.npz.TRUNK_KEYS + train_value.npz.HEAD_KEYS[1])
    tr.close()
