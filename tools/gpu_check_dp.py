"""2+ ranks: three data-parallel update steps; replicas must stay bit-identical, and the overlapped all-reduce
(W_ih block on a second stream) must give the same parameters as the single all-reduce."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
from cadre_b200 import fixtures as R
from cadre_b200.learner import Learner, RolloutPool
W = 2
results = {}
for mode in ("overlap", "single"):
    os.environ["CADRE_NO_ALLREDUCE_OVERLAP"] = "0" if mode == "overlap" else "1"
    learner = Learner(W, 100, R.ppo_fixture_state(0), dev, seeds=[rank * W + w for w in range(W)])
    pool = RolloutPool(W, dict(num_steps=200, mini_batch_num=2, feature_dims=530, seq_length=8, use_gae=True, gamma=0.99, tau=0.95), dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
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
    for k in range(3):
        learner.update_step(pool.storages, idx[k % len(idx)])
    torch.cuda.synchronize()
    p = learner.params.clone()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    results[mode] = p
    if rank == 0:
        print(f"{mode}: replicas bit-identical across {world} ranks: {same}; |theta| = {p.norm().item():.6f}", flush=True)
    assert same
d = (results["overlap"] - results["single"]).abs().max().item()
if rank == 0:
    print(f"overlapped vs single all-reduce: max |delta theta| = {d:.3e}", flush=True)
assert d < 1e-6
dist.barrier(); dist.destroy_process_group()
