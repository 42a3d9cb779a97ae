"""In-situ clock stamps of the LSTM-step GEMM (EPI_LSTM) inside a real PPO update (W=4, mb=100)."""
import os, sys, statistics
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = "cuda:0"
torch.zeros(1, device=dev)
clk = torch.zeros(4 * 17 * 8 * 8 * 4, dtype=torch.int64, device=dev)
os.environ["CADRE_DBG_CLK_EPI"] = f"{sys.argv[1] if len(sys.argv) > 1 else 1}:{clk.data_ptr()}"
from cadre_b200.learner import Learner, RolloutPool
from cadre_b200 import fixtures as R
W = 4
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
for k in range(4):
    learner.update_step(pool.storages, idx[k % len(idx)])
torch.cuda.synchronize()
c = clk.view(-1, 8).cpu()
valid = c[(c[:, 5] != 0)]
print("CTAs with an epilogue stamp:", valid.shape[0], "of", int((c[:, 0] != 0).sum()), "started")
d = valid - valid[:, :1]
names = ["start", "after alloc+sync", "first full (MMA)", "MMA all issued", "epilogue sees tmem_full", "epilogue done", "after dealloc", "phase 1 done"]
for i, n in enumerate(names):
    col = d[:, i].tolist()
    print(f"{n:28s} median {statistics.median(col):9.0f} cyc  min {min(col):9.0f} max {max(col):9.0f}")
allc = c[c[:, 0] != 0]
print("kernel span (first start -> last stamp):", int(allc[:, 6].max() - allc[:, 0].min()), "cycles (SM clocks are not synchronised across SMs: indicative)")
