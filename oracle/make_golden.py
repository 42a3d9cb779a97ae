"""ORACLE — TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic inputs. Run in the build container:

    python -m oracle.make_golden

The reference has no golden vectors of its own; these files are the pin. Inputs and weights are
regenerated from seeds by oracle/restate.py (`synthetic_*`, `*_fixture_state`), so only outputs are stored.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, restate as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
THREADS = 8  # goldens are generated and re-checked with this thread count


def tensor_stats(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.norm().item(), t.abs().max().item()], dtype=np.float64)


def golden_encoder():
    out = {}
    for tag, peaky in (("base", False), ("peaky", True)):
        sd = R.danet_fixture_state(seed=0, peaky=peaky)
        net, _ = ref_shim.build_reference_danet(sd)
        rs = np.random.RandomState(1000)
        tick = R.synthetic_tick(rs)
        x = torch.from_numpy(R.pre_process(tick["rgb"], tick["route_fig"].copy()))
        with torch.no_grad():
            l4 = net.backbone(x)
            da = net.da_head(l4)
            lat = net.get_latent_feature(x, "concate")
        out[f"{tag}_latent"] = lat.numpy()
        out[f"{tag}_l4_stats"] = tensor_stats(l4)
        out[f"{tag}_da_stats"] = tensor_stats(da)
        out[f"{tag}_l4_slice"] = l4[:, :8, :, :].numpy()
    np.savez_compressed(os.path.join(OUT, "encoder.npz"), **out)
    print("encoder.npz", {k: v.shape for k, v in out.items()})


def golden_agent_feature():
    """CadreAgent.get_latent_feature (agent.py:97-112) incl. pre_process quirks, through the reference agent."""
    agent, _ = ref_shim.build_reference_agent(R.danet_fixture_state(0), R.ppo_fixture_state(0))
    rs = np.random.RandomState(2000)
    tick = R.synthetic_tick(rs)
    tick["route_fig"][3] = (rs.rand(256, 144) * 200).astype(np.uint8)  # non-binary map: exercises truncation
    tick["route_fig"][5] = 0                                            # all-zero map: max == 0 branch
    with torch.no_grad():
        feat = agent.get_latent_feature({k: (v.copy() if hasattr(v, "copy") else v) for k, v in tick.items()})
        # get_value on that feature for every command (agent.py:143-164)
        vals = []
        for c in range(4):
            vs, vt = agent.get_value(False, (feat, c), (feat, c))
            vals.append([vs.item(), vt.item()])
    np.savez_compressed(os.path.join(OUT, "agent_feature.npz"), feature=feat.numpy(),
                        values=np.array(vals, dtype=np.float32))
    print("agent_feature.npz", feat.shape)


def golden_gae():
    out = {}
    for i, (T, seed) in enumerate(((200, 0), (200, 1), (800, 2), (7, 3))):
        rs = np.random.RandomState(seed)
        st = R.synthetic_storage(rs, T=T, feature_dims=8, seq=1)
        if i == 1:
            st["masks"][::5] = 0.0
        ref = ref_shim.reference_storage(st, dict(num_steps=T, mini_batch_num=2, feature_dims=8, seq_length=1,
                                                  use_gae=True, gamma=R.GAMMA, tau=R.TAU))
        next_value = torch.tensor([[0.37 * (i + 1)]])
        ref.compute_returns(next_value)
        adv = ref.returns[:-1] - ref.value_preds[:-1]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)   # train.py:82-88
        out[f"returns_{i}"] = ref.returns.numpy()
        out[f"adv_{i}"] = adv.numpy()
        out[f"meta_{i}"] = np.array([T, seed, next_value.item()], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)
    print("gae.npz")


def golden_indices():
    out = {}
    for seed in (0, 1, 7):
        rs = np.random.RandomState(seed)
        st_s = ref_shim.reference_storage(R.synthetic_storage(rs, feature_dims=8, seq=1),
                                          dict(num_steps=200, mini_batch_num=2, feature_dims=8, seq_length=1,
                                               use_gae=True, gamma=R.GAMMA, tau=R.TAU))
        st_t = ref_shim.reference_storage(R.synthetic_storage(rs, feature_dims=8, seq=1),
                                          dict(num_steps=200, mini_batch_num=2, feature_dims=8, seq_length=1,
                                               use_gae=True, gamma=R.GAMMA, tau=R.TAU))
        torch.manual_seed(seed)
        adv = torch.zeros(200, 1)
        rows = []
        # recover the yielded indices through the `action` column, which we set to arange
        st_s.action[:200, 0] = torch.arange(200)
        st_t.action[:200, 0] = torch.arange(200)
        for _ in range(R.PPO_EPOCH):  # train.py:93-96: zip(steer generator, throttle generator)
            for s_samp, t_samp in zip(st_s.feed_forward_generator(adv), st_t.feed_forward_generator(adv)):
                rows.append(np.stack([s_samp[1][:, 0].numpy(), t_samp[1][:, 0].numpy()]))
        out[f"idx_{seed}"] = np.stack(rows).astype(np.int64)  # [8 steps, 2 heads, 100]
    np.savez_compressed(os.path.join(OUT, "indices.npz"), **out)
    print("indices.npz", out["idx_0"].shape)


def golden_update(workers=2):
    """W workers x (update_policy -> add_gradient), then the chief body (chief.py:13-22), all in-process."""
    torch.set_num_threads(THREADS)
    ppo_sd = R.ppo_fixture_state(0)
    agent, _ = ref_shim.build_reference_agent(R.danet_fixture_state(0), ppo_sd)
    from ppo_agent.models import Shared_grad_buffers, create_model
    cfg = ref_shim.load_agent_config()
    _, shared = create_model(cfg.agent_cfg.model_cfg, load_vae=False)
    params = []
    for name in shared:
        shared[name].load_state_dict(ppo_sd[name])
        params += list(shared[name].parameters())
    opt = torch.optim.Adam(params, lr=cfg.train_cfg.lr)
    bufs = Shared_grad_buffers(shared, torch.device("cpu"))
    assert list(shared.keys()) == R.PPO_MODULE_ORDER, list(shared.keys())

    out = {}
    losses = []
    for w in range(workers):
        rs = np.random.RandomState(100 + w)
        st_s = R.synthetic_storage(rs, actions=R.STEER_ACTIONS)
        st_t = R.synthetic_storage(rs, actions=R.THROTTLE_ACTIONS)
        # non-zero recurrent state in one worker: kernels must honour arbitrary h0/c0
        if w == 1:
            for st in (st_s, st_t):
                st["hn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
                st["cn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
        rs_s, rs_t = ref_shim.reference_storage(st_s), ref_shim.reference_storage(st_t)
        rs_s.compute_returns(torch.tensor([[0.1]]))
        rs_t.compute_returns(torch.tensor([[-0.2]]))
        advs = []
        for r_ in (rs_s, rs_t):
            a = r_.returns[:-1] - r_.value_preds[:-1]
            advs.append((a - a.mean()) / (a.std() + 1e-8))
        torch.manual_seed(500 + w)
        s_samp = next(iter(rs_s.feed_forward_generator(advs[0])))
        t_samp = next(iter(rs_t.feed_forward_generator(advs[1])))
        agent.update_model(shared)
        vl, al, el = agent.update_policy(s_samp, t_samp)
        losses.append([vl, al, el])
        bufs.add_gradient(agent.model_dict)
        if w == 0:
            gstats, gslices = [], []
            for name in R.PPO_MODULE_ORDER:
                for pn, p in agent.model_dict[name].named_parameters():
                    gstats.append(tensor_stats(p.grad))
                    gslices.append(p.grad.flatten()[:32].numpy().copy())
            out["w0_grad_stats"] = np.stack(gstats)
            out["w0_grad_slices"] = np.stack([np.pad(g, (0, 32 - len(g))) for g in gslices])
    out["losses"] = np.array(losses, dtype=np.float64)

    # chief body, chief.py:13-21
    opt.zero_grad()
    norms = []
    for name in shared:
        for n, p in shared[name].named_parameters():
            p._grad = bufs.grads[name + "_" + n + "_grad"].clone().detach()
        norms.append(float(torch.nn.utils.clip_grad_norm_(shared[name].parameters(), cfg.train_cfg.max_grad_norm)))
    before = {name: {n: p.detach().clone() for n, p in shared[name].named_parameters()} for name in shared}
    opt.step()
    pstats, dstats, pslices = [], [], []
    for name in R.PPO_MODULE_ORDER:
        for n, p in shared[name].named_parameters():
            pstats.append(tensor_stats(p))
            dstats.append(tensor_stats(p.detach() - before[name][n]))
            pslices.append(np.pad(p.detach().flatten()[:32].numpy(), (0, max(0, 32 - p.numel()))))
    out["module_grad_norms"] = np.array(norms)
    out["post_param_stats"] = np.stack(pstats)
    out["post_delta_stats"] = np.stack(dstats)
    out["post_param_slices"] = np.stack(pslices)
    np.savez_compressed(os.path.join(OUT, "update.npz"), **out)
    print("update.npz losses", out["losses"], "module norms", out["module_grad_norms"][:4])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(THREADS)
    golden_gae()
    golden_indices()
    golden_encoder()
    golden_agent_feature()
    golden_update()


if __name__ == "__main__":
    main()
