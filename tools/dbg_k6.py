import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_reinforce_gpu as T
from iago_b200.train_rl import ReinforceTrainer, N_PARAMS
from iago_b200 import npz
from oracle import reinforce_ref, nets
path='/root/repo/baseline/_ref/models/RL/model2.npz'
states, actions, rewards = T.golden_batch()
own, opp = T.to_device(states)
a = torch.from_numpy(actions.astype(np.int8)).cuda(); r = torch.from_numpy(rewards).cuda()
g={}
for tc in (False, True):
    tr = ReinforceTrainer(path, max_positions=256, tensor_cores=tc, slot=4)
    tr.gradient(own, opp, a, r); torch.cuda.synchronize()
    g[tc]=tr.grad.cpu().numpy().astype(np.float64); tr.close()
total, ref, pred = reinforce_ref.loss_and_grad(nets.load_params(path, np.float64), states, actions, rewards)
shapes = {k: v.shape for k, v in npz.unflatten(np.zeros(N_PARAMS, np.float32), 0).items()}
o=0
for k in reinforce_ref.KEYS:
    n=int(np.prod(shapes[k])); R=ref[k].reshape(-1); s=np.abs(R).max()
    print(f"{k:18s} scale {s:9.3e}  simt-ref {np.abs(g[False][o:o+n]-R).max()/s:8.2e}  tc-ref {np.abs(g[True][o:o+n]-R).max()/s:8.2e}  tc-simt {np.abs(g[True][o:o+n]-g[False][o:o+n]).max()/s:8.2e}")
    o+=n
