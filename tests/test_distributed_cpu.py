"""World-size-2 gloo test of the data-parallel contract on CPU: per-rank worker RNG streams are bit-exact and
all_reduce(SUM) of per-rank flat gradients followed by the same deterministic clip + Adam (oracle maths here, the
CUDA kernels in -m gpu tests) leaves both ranks with identical parameters equal to the reference chief step on
the summed gradient (chief.py:13-21)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cadre_b200 import ppo_params as P
    from cadre_b200.learner import Learner
    from cadre_b200.storage import RolloutStorage
    from oracle import restate as R
    torch.set_num_threads(2)
    # (1) index streams: rank r owns workers 2r, 2r+1 seeded 500 + id (same as separate reference processes)
    L = Learner.__new__(Learner)
    L.mini_batch, L._rng_states = 100, []
    keep = torch.get_rng_state()
    for w in range(2):
        torch.manual_seed(500 + 2 * rank + w)
        L._rng_states.append(torch.get_rng_state())
    torch.set_rng_state(keep)
    sts = [tuple(RolloutStorage(200, 2, 530, 8, 530, True, 0.99, 0.95) for _ in range(2)) for _ in range(2)]
    idx = L.sample_epoch_indices(sts)
    ok_idx = True
    for w in range(2):
        torch.manual_seed(500 + 2 * rank + w)
        ok_idx &= list(idx[0, w, 0]) == R.minibatch_indices()[0]
    # (2) gradient exchange on a small synthetic "gradient": sum over ranks, then identical clip + Adam
    g = torch.Generator().manual_seed(1000 + rank)
    sd = R.ppo_fixture_state(0)
    local = {m: {n: torch.randn(t.shape, generator=g) * 0.01 for n, t in d.items()} for m, d in sd.items()}
    flat = P.pack_state(local)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    summed = P.unpack_state(flat)
    params = {m: {n: t.clone() for n, t in d.items()} for m, d in sd.items()}
    adam = {m: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
            for m, d in sd.items()}
    R.chief_step(params, summed, adam, step=1)
    post = P.pack_state(params)
    gathered = [torch.zeros_like(post) for _ in range(world)]
    dist.all_gather(gathered, post)
    identical = all(torch.equal(gathered[0], x) for x in gathered)
    # the sum really is the sum of both ranks' gradients
    g0 = torch.Generator().manual_seed(1000)
    g1 = torch.Generator().manual_seed(1001)
    m0 = "steer_ppo_0"
    n0 = "control.linear.0.weight"
    e0 = torch.randn(sd[m0][n0].shape, generator=g0) * 0.01
    e1 = torch.randn(sd[m0][n0].shape, generator=g1) * 0.01
    ok_sum = torch.allclose(summed[m0][n0], e0 + e1, atol=1e-7)
    if rank == 0:
        torch.save({"ok_idx": ok_idx, "identical": identical, "ok_sum": ok_sum}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_contract(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["ok_idx"] and res["identical"] and res["ok_sum"]
