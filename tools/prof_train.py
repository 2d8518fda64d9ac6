import cProfile, pstats, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from iago_b200 import network
from iago_b200.train_rl import ReinforceTrainer
mdir = "baseline/_ref/models"
opp = network.SLPolicy().load(os.path.join(mdir, "RL", "model0.npz"))
tr = ReinforceTrainer(os.path.join(mdir, "rl_model.npz"), max_positions=8192)
tr.train_set(opp, n_games=2048, seed=1)
torch.cuda.synchronize()
for _ in range(2):
    t0=time.perf_counter(); d = tr.play_set(opp, 2048, seed=3); torch.cuda.synchronize(); t1=time.perf_counter()
    tr.gradient(d["own"], d["opp"], d["action"], d["reward"]); torch.cuda.synchronize(); t2=time.perf_counter()
    tr.update(); torch.cuda.synchronize(); t3=time.perf_counter()
    print(f"play_set {1e3*(t1-t0):.1f} ms  gradient {1e3*(t2-t1):.1f} ms  update {1e3*(t3-t2):.1f} ms  positions {d['own'].numel()}")
pr = cProfile.Profile(); pr.enable()
tr.train_set(opp, n_games=2048, seed=2); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
