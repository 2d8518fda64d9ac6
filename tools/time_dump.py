import os, sys, torch, numpy as np
sys.path.insert(0, os.getcwd())
import iago_b200
from iago_b200 import boards
eng = iago_b200.default_engine(0)
eng.load_net(0, "baseline/_ref/models/sl_model.npz")
n = 8192
rs = np.random.RandomState(0)
p1 = torch.from_numpy(rs.randint(0, 2**62, size=n, dtype=np.int64)).cuda()
p2 = torch.from_numpy(rs.randint(0, 2**62, size=n, dtype=np.int64)).cuda() & ~p1
col = torch.ones(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("plain forward  ms", t(lambda: eng.policy_forward(0, p1, p2, col, probs=False)))
print("forward + acts ms", t(lambda: eng.policy_forward_acts(0, p1, p2, col)))
