"""Per-step cycle breakdown of the persistent LSTM recurrence kernels (lstm_seq.cuh) inside a real PPO update
(W=4, mb=100): clock64 stamps written by one thread per role (cadre_debug_clk). Per CTA and step t:
  0 producer enters step | 1 hand-off counter reached | 2 last TMA load issued | 3 last MMA committed |
  4 epilogue sees the accumulator | 5 exchange slice stored | 6 slice published | 7 remaining stores issued
and per CTA: 64 kernel start | 65 weights resident | 66 griddepcontrol.wait returned."""
import ctypes, os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
from cadre_b200.learner import Learner, RolloutPool
from cadre_b200 import fixtures as R
dev = "cuda:0"
W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MB = int(sys.argv[2]) if len(sys.argv) > 2 else 100
T = 2 * MB
learner = Learner(W, MB, R.ppo_fixture_state(0), dev, seeds=list(range(W)))
pool = RolloutPool(W, dict(num_steps=T, mini_batch_num=2, feature_dims=530, seq_length=8, use_gae=True, gamma=0.99, tau=0.95), dev)
g = torch.Generator(device=dev).manual_seed(0)
b = pool.batched
b["obs"].copy_(torch.randn(b["obs"].shape, device=dev, generator=g))
b["rewards"].copy_(torch.rand(b["rewards"].shape, device=dev, generator=g))
b["masks"].fill_(1.0)
b["command"].copy_(torch.randint(0, 4, b["command"].shape, device=dev, generator=g, dtype=torch.int32))
b["action_log_probs"].fill_(-1.5)
b["value_preds"].copy_(torch.randn(b["value_preds"].shape, device=dev, generator=g))
for w in range(W):
    b["action"][2 * w].copy_(torch.randint(0, 33, (T + 1, 1), device=dev, generator=g))
    b["action"][2 * w + 1].copy_(torch.randint(0, 3, (T + 1, 1), device=dev, generator=g))
pool.compute_returns(torch.zeros(W, 2, device=dev))
idx = learner.sample_epoch_indices(pool.storages)
for k in range(3):
    learner.update_step(pool.storages, idx[k % len(idx)])
torch.cuda.synchronize()
clk = torch.zeros(2 * 136 * 72, dtype=torch.int64, device=dev)
L = _lib.lib()
L.cadre_debug_clk(ctypes.c_void_p(clk.data_ptr()))
learner.update_step(pool.storages, idx[1])
torch.cuda.synchronize()
L.cadre_debug_clk(ctypes.c_void_p(0))
learner.engine.check()
c = clk.view(2, 136, 9, 8).cpu()
MHZ = 1.9e3   # ~cycles per microsecond (SM clock under load)


def med(x):
    x = [v for v in x if v == v]
    return statistics.median(x) if x else float("nan")


for d, name, steps in ((0, "forward", range(0, 8)), (1, "backward", range(7, 0, -1))):
    k = c[d]
    started = k[:, 8, 0] != 0
    print(f"== {name}: {int(started.sum())} CTAs")
    kk = k[started].double()
    print(f"  weights resident after {med((kk[:, 8, 1] - kk[:, 8, 0]).tolist()) / MHZ:7.2f} us; "
          f"pdl wait returns after {med((kk[:, 8, 2] - kk[:, 8, 0]).tolist()) / MHZ:7.2f} us")
    first = steps[0]
    end_stamp = 6
    total = (kk[:, steps[-1], end_stamp if d == 0 else 5] - kk[:, 8, 2])
    print(f"  pdl-wait -> last publish: median {med(total.tolist()) / MHZ:7.2f} us")
    for t in steps:
        s = kk[:, t, :]
        row = {
            "wait hand-off": s[:, 1] - s[:, 0],
            "TMA issue": s[:, 2] - s[:, 1],
            "MMA done after last TMA": s[:, 3] - s[:, 2],
            "epilogue sees acc after MMA commit": s[:, 4] - s[:, 3],
            "epilogue compute+stores": s[:, 5] - s[:, 4],
            "fence+publish": s[:, 6] - s[:, 5],
            "bulk stores": s[:, 7] - s[:, 6],
            "step (producer enter -> publish)": s[:, 6] - s[:, 0],
        }
        print(f"  t={t}: " + " | ".join(f"{n} {med(v.tolist()) / MHZ:6.2f}" for n, v in row.items()))
