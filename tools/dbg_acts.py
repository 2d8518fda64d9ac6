import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_reinforce_gpu as T
import iago_b200
from iago_b200 import boards
from oracle import nets
path='/root/repo/baseline/_ref/models/RL/model2.npz'
states, actions, rewards = T.golden_batch()
own, opp = T.to_device(states)
eng = iago_b200.default_engine(0)
eng.load_net(0, path)
col = torch.ones(own.numel(), dtype=torch.uint8, device='cuda')
logits, acts = eng.policy_forward_acts(0, own, opp, col)
torch.cuda.synchronize()
p = nets.load_params(path, np.float64)
s = np.asarray(states).reshape(-1,8,8)
x = np.stack([s==1, s==2], axis=1).astype(np.float64)
h = x
for l in range(8):
    h = np.maximum(nets.conv2d(h, p[f"block{l+1}/conv/W"], p[f"block{l+1}/conv/b"]), 0)
    a = acts[l].cpu().numpy()
    d = np.abs(a - h)
    flips = ((a > 0) != (h > 0))
    # magnitude of reference values where masks differ
    print(f"block{l+1}: max|act| {h.max():8.3f} max err {d.max():.2e}  mask flips {flips.sum()} of {flips.size}  max |ref| at flips {np.abs(h[flips]).max() if flips.any() else 0:.2e} max |gpu| at flips {np.abs(a[flips]).max() if flips.any() else 0:.2e}")
