"""Rollout ingest: pinned host frames -> RolloutStorage.obs through cadre_b200.ingest.RolloutIngest.
The unique-frame path (each distinct frame shipped + encoded once, windows assembled by cadre_window_scatter) must give
exactly the observations of the reference's per-tick act() loop (agent.py:97-112 + train.py:66-72: the full 8-frame
window encoded every tick and inserted into both heads' storages)."""
import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def _rollout_frames(W, K, seed):
    g = torch.Generator().manual_seed(seed)
    rgb = torch.randint(0, 256, (W, K, 144, 256, 3), dtype=torch.uint8, generator=g)
    route = (torch.rand(W, K, 256, 144, generator=g) < 0.1).to(torch.uint8).mul(255)
    route[0, 3] = (torch.rand(256, 144, generator=g) * 200).to(torch.uint8)     # non-binary map (truncation quirk)
    route[W - 1, 5] = 0                                                        # all-zero map
    meas = torch.rand(W, K, 3, dtype=torch.float64, generator=g)
    return rgb, route, meas


def _pin(t):
    return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)


@pytest.mark.parametrize("max_chunk,streams", [(64, 2), (640, 1)])
def test_unique_frame_ingest_equals_per_tick_window_encoding(max_chunk, streams):
    from cadre_b200.ingest import RolloutIngest
    from cadre_b200.learner import RolloutPool
    W, T, S = 3, 21, 8
    K = T + S - 1
    sd = R.danet_fixture_state(0)
    ing = RolloutIngest(sd, "cuda:0", W, T, S, 530, max_chunk=max_chunk, streams=streams)
    cfg = dict(num_steps=T, mini_batch_num=1, feature_dims=530, seq_length=S, use_gae=True, gamma=0.99, tau=0.95)
    pools = [RolloutPool(W, cfg, "cuda:0") for _ in range(3)]
    for p in pools:
        p.batched["obs"].fill_(-5.0)
    rgb, route, meas = _rollout_frames(W, K, seed=3)
    win = (torch.arange(T).view(T, 1) + torch.arange(S).view(1, S)).flatten()
    rgb_w = rgb[:, win].reshape(W, T, S, 144, 256, 3)
    route_w = route[:, win].reshape(W, T, S, 256, 144)
    meas_w = meas[:, win].reshape(W, T, S, 3)

    # (a) product path: pinned host, distinct frames only
    ing.encode(_pin(rgb), _pin(route), _pin(meas), pools[0].batched["obs"], unique=True)
    assert ing.frames_encoded_last == W * K
    assert ing.h2d_bytes_last == W * K * (144 * 256 * 3 + 256 * 144 + 24)
    # (b) every window frame, pinned host (what the reference's act() loop moves and encodes)
    ing.encode(_pin(rgb_w), _pin(route_w), _pin(meas_w), pools[1].batched["obs"], unique=False)
    assert ing.frames_encoded_last == W * T * S
    # (c) every window frame, inputs already on the device (bench.py's resident leg)
    ing.encode(rgb_w.cuda(), route_w.cuda(), meas_w.cuda(), pools[2].batched["obs"], unique=False)
    torch.cuda.synchronize()
    a, b, c = (p.batched["obs"] for p in pools)
    assert torch.equal(a, b) and torch.equal(b, c)
    assert float(a[:, T].min()) == -5.0 and float(a[:, T].max()) == -5.0        # slot T is not a rollout step
    assert torch.equal(a[0::2], a[1::2])                                        # steer and throttle storages agree
    # against the oracle, per tick like the reference: three ticks of two workers
    torch.set_num_threads(8)
    for w, t in ((0, 0), (0, 2), (1, 4), (2, T - 1)):
        with torch.no_grad():
            ref = R.agent_latent_feature(rgb_w[w, t].numpy(), route_w[w, t].numpy().copy(), meas_w[w, t].numpy(), sd)
        got = pools[0].storages[w][1].obs[t].cpu()
        assert torch.equal(got[:, 512:], ref[:, 512:])
        err = ((got[:, :512] - ref[:, :512]).double().norm() / ref[:, :512].double().norm()).item()
        assert err < 1e-2, (w, t, err)


def test_prefetched_rollout_equals_in_step_staging():
    """RolloutIngest.prefetch (the next rollout's frames cross PCIe under the current update phase) feeds encode() the
    same bytes as the chunked in-step staging: identical observations, for two consecutive rollouts, and a rollout that
    was NOT the prefetched one falls back to staging."""
    from cadre_b200.ingest import RolloutIngest
    from cadre_b200.learner import RolloutPool
    W, T, S = 2, 13, 8
    K = T + S - 1
    ing = RolloutIngest(R.danet_fixture_state(0), "cuda:0", W, T, S, 530, max_chunk=32, streams=2)
    cfg = dict(num_steps=T, mini_batch_num=1, feature_dims=530, seq_length=S, use_gae=True, gamma=0.99, tau=0.95)
    pa, pb = RolloutPool(W, cfg, "cuda:0"), RolloutPool(W, cfg, "cuda:0")
    rollouts = [tuple(_pin(t) for t in _rollout_frames(W, K, seed=s)) for s in (11, 12)]
    nbytes = W * K * (144 * 256 * 3 + 256 * 144 + 24)
    for i, host in enumerate(rollouts):
        ing.encode(*host, pa.batched["obs"], unique=True)                   # in-step staging
        assert ing.h2d_bytes_last == nbytes
        ing.prefetch(*host)
        torch.cuda._sleep(2_000_000)                                        # something else runs meanwhile
        ing.encode(*host, pb.batched["obs"], unique=True)                   # consumes the prefetched copy
        assert ing.h2d_bytes_last == nbytes and ing._landing_key is None
        torch.cuda.synchronize()
        assert torch.equal(pa.batched["obs"], pb.batched["obs"]), i
    # prefetched rollout 1, but rollout 0 is encoded: staging path, and the prefetched copy stays valid for later
    ing.prefetch(*rollouts[1])
    ing.encode(*rollouts[0], pa.batched["obs"], unique=True)
    assert ing._landing_key is not None
    ing.encode(*rollouts[1], pb.batched["obs"], unique=True)
    assert ing._landing_key is None
    torch.cuda.synchronize()
    assert not torch.equal(pa.batched["obs"], pb.batched["obs"])
    ing.encode(*rollouts[1], pa.batched["obs"], unique=True)
    torch.cuda.synchronize()
    assert torch.equal(pa.batched["obs"], pb.batched["obs"])
    with pytest.raises(Exception, match="pinned"):
        ing.prefetch(*[t.clone() for t in _rollout_frames(W, K, seed=1)])


def test_ingest_rejects_pageable_host_memory_and_wrong_shapes():
    from cadre_b200._lib import CadreError
    from cadre_b200.ingest import RolloutIngest
    ing = RolloutIngest(R.danet_fixture_state(0), "cuda:0", 1, 4, 8, 530, max_chunk=16, streams=1)
    obs = torch.zeros(2, 5, 8, 530, device="cuda")
    rgb, route, meas = _rollout_frames(1, 11, seed=1)
    with pytest.raises(CadreError, match="pinned"):
        ing.encode(rgb, route, meas, obs)
    with pytest.raises(CadreError, match="leading shape"):
        ing.encode(_pin(rgb[:, :10]), _pin(route[:, :10]), _pin(meas[:, :10]), obs)
