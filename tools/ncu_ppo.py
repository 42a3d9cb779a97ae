"""Short PPO run for ncu (--profile-from-start off): W=4 workers x mb=100, one warm-up update step, one profiled."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200.learner import Learner, RolloutPool
from cadre_b200 import fixtures as R
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda:0"
learner = Learner(W, 100, R.ppo_fixture_state(0), dev, seeds=list(range(W)))
pool = RolloutPool(W, dict(num_steps=200, mini_batch_num=2, feature_dims=530, seq_length=8, use_gae=True, gamma=0.99, tau=0.95), dev)
g = torch.Generator(device=dev).manual_seed(0)
b = pool.batched
b["obs"].copy_(torch.randn(b["obs"].shape, device=dev, generator=g))
b["rewards"].copy_(torch.rand(b["rewards"].shape, device=dev, generator=g))
b["masks"].fill_(1.0)
b["command"].copy_(torch.randint(0, 4, b["command"].shape, device=dev, generator=g, dtype=torch.int32))
b["action_log_probs"].fill_(-1.5)
b["value_preds"].copy_(torch.randn(b["value_preds"].shape, device=dev, generator=g))
for w in range(W):
    b["action"][2 * w].copy_(torch.randint(0, 33, (201, 1), device=dev, generator=g))
    b["action"][2 * w + 1].copy_(torch.randint(0, 3, (201, 1), device=dev, generator=g))
pool.compute_returns(torch.zeros(W, 2, device=dev))
idx = learner.sample_epoch_indices(pool.storages)
learner.update_step(pool.storages, idx[0])          # warm-up (also builds the TMA descriptors)
torch.cuda.synchronize()
torch.cuda.profiler.start()                         # ncu --profile-from-start off: exactly one update step
learner.update_step(pool.storages, idx[1 % len(idx)])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
