"""Drop-in API on the B200: CadreAgent / RolloutStorage / Learner behave like the reference objects
(ppo_agent/agent.py, storage.py, train.py + chief.py) and agree with the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


def rel(got, ref):
    got, ref = got.double().flatten().cpu(), ref.double().flatten().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def agent():
    from cadre_b200.agent import CadreAgent
    from cadre_b200.config import load_config
    cfg = load_config()
    return CadreAgent(**cfg.agent_cfg, danet_state=R.danet_fixture_state(0), ppo_state=R.ppo_fixture_state(0),
                      max_encoder_batch=8)


def test_act_and_get_value(agent):
    torch.set_num_threads(8)
    tick = R.synthetic_tick(np.random.RandomState(3))
    torch.manual_seed(11)
    feat, action, logp, value, hidden = agent.act({k: (v.copy() if hasattr(v, "copy") else v) for k, v in tick.items()})
    assert feat.shape == (8, 530) and feat.dtype == torch.float32
    assert action[0].dim() == 0 and action[0].dtype == torch.int64 and 0 <= int(action[0]) < 33
    assert 0 <= int(action[1]) < 3 and logp[0].shape == (1, 1) and value[1].shape == (1, 1)
    assert hidden[0].shape == (1, 530) and float(hidden[0].abs().sum()) == 0.0     # never updated (quirk 3)
    # oracle: same feature -> LSTM of the command -> critic / actor
    sd, ppo = R.danet_fixture_state(0), R.ppo_fixture_state(0)
    with torch.no_grad():
        ref_feat = R.agent_latent_feature(tick["rgb"], tick["route_fig"].copy(), tick["measurements"], sd)
        assert rel(feat, ref_feat) < 1e-2
        c = tick["command"]
        h0 = torch.zeros(1, 530)
        for h, head in enumerate(("steer", "throttle")):
            f, _ = R.lstm_forward(ref_feat, h0, h0, ppo[f"{head}_lstm_{c}"])
            v, lp, _ = R.evaluate_actions(f, action[h].reshape(1, 1), ppo[f"{head}_ppo_{c}"])
            assert abs(value[h].item() - v.item()) < 1e-2 * max(1.0, abs(v.item()))
            assert abs(logp[h].item() - lp.item()) < 1e-2
    vs, vt = agent.get_value(False, (feat, c), (feat, (c + 1) % 4))
    assert vs.shape == (1, 1) and abs(vs.item() - value[0].item()) < 1e-5
    z = agent.get_value(True, (feat, c), (feat, c))
    assert float(z[0]) == 0.0 and z[0].shape == (1,)
    ctrl = agent.convert_action(action)
    assert len(ctrl) == 3 and -1.0 <= ctrl[0] <= 1.0


def test_storage_and_update_policy_like_train_loop(agent):
    from cadre_b200.storage import RolloutStorage
    torch.set_num_threads(8)
    rs = np.random.RandomState(42)
    cpu = [R.synthetic_storage(rs, actions=a) for a in (R.STEER_ACTIONS, R.THROTTLE_ACTIONS)]
    sts = []
    for c in cpu:
        st = RolloutStorage(num_steps=200, mini_batch_num=2, feature_dims=530, seq_length=8, hidden_size=530,
                            use_gae=True, gamma=0.99, tau=0.95)
        for k in ("obs", "rewards", "value_preds", "action_log_probs", "action", "masks", "command"):
            getattr(st, k).copy_(c[k])
        st.to(agent.device)
        st.compute_returns(torch.tensor([[0.3]]))
        ret, vp = R.compute_returns(c["rewards"], c["value_preds"], c["masks"], torch.tensor([[0.3]]))
        assert rel(st.returns[:200], ret[:200]) < 1e-5
        assert rel(st.advantages, R.normalized_advantages(ret, vp)) < 1e-5
        c["returns"], c["value_preds"] = ret, vp
        sts.append(st)
    # identical index stream to the reference sampler for the same torch seed
    torch.manual_seed(5)
    gens = [s.feed_forward_generator() for s in sts]
    mbs = [next(g) for g in gens]
    torch.manual_seed(5)
    assert mbs[0].indices == R.minibatch_indices()[0] and mbs[1].indices == R.minibatch_indices()[0]
    assert len(tuple(mbs[0])) == 9 and tuple(mbs[0])[0].shape == (800, 530)
    losses = agent.update_policy(mbs[0], mbs[1])
    ppo = R.ppo_fixture_state(0)
    params = {m: {n: t.clone().requires_grad_(True) for n, t in d.items()} for m, d in ppo.items()}
    advs = [R.normalized_advantages(c["returns"], c["value_preds"]) for c in cpu]
    ref = R.update_policy(R.gather_minibatch(cpu[0], advs[0], mbs[0].indices),
                          R.gather_minibatch(cpu[1], advs[1], mbs[1].indices), params)
    np.testing.assert_allclose(np.array(losses), np.array(ref), rtol=1e-3)
    # the reference-style tuple input takes the same path
    losses2 = agent.update_policy(tuple(mbs[0]), tuple(mbs[1]))
    np.testing.assert_allclose(np.array(losses2), np.array(losses), rtol=1e-5)
    g = agent.model_dict["steer_lstm_0"].named_parameters()
    name, p = next(iter(g))
    assert name == "rnn.weight_ih" and rel(p.grad, params["steer_lstm_0"]["rnn.weight_ih"].grad) < 5e-2


def test_learner_two_steps_match_chief_contract():
    """W=2 workers batched on one GPU == sum of per-worker gradients -> per-module clip -> Adam (chief.py)."""
    from cadre_b200.learner import Learner, RolloutPool
    torch.set_num_threads(8)
    ppo = R.ppo_fixture_state(0)
    learner = Learner(2, 100, ppo, "cuda:0", seeds=[500, 501])
    pool = RolloutPool(2, dict(num_steps=200, mini_batch_num=2, feature_dims=530, seq_length=8, use_gae=True,
                               gamma=0.99, tau=0.95), "cuda:0")
    cpu = []
    for w in range(2):
        rs = np.random.RandomState(100 + w)
        pair = [R.synthetic_storage(rs, actions=a) for a in (R.STEER_ACTIONS, R.THROTTLE_ACTIONS)]
        for h, c in enumerate(pair):
            for k in ("obs", "rewards", "value_preds", "action_log_probs", "action", "masks", "command"):
                getattr(pool.storages[w][h], k).copy_(c[k])
        cpu.append(pair)
    nv = torch.tensor([[0.1, -0.2], [0.1, -0.2]])
    pool.compute_returns(nv)
    idx = learner.sample_epoch_indices(pool.storages)          # [2 minibatches, W, 2, 100]
    assert idx.shape == (2, 2, 2, 100)
    for w in range(2):                                         # per-worker RNG streams == torch.manual_seed(500+w)
        torch.manual_seed(500 + w)
        assert list(idx[0, w, 0]) == R.minibatch_indices()[0]
    before = learner.params.clone()
    learner.update_step(pool.storages, idx[0])
    assert learner.step_count == 1 and not torch.equal(before, learner.params)
    L = learner.scaled_losses()
    assert L.shape == (2, 3) and torch.isfinite(L).all()
    # oracle emulation of the same step
    params = {m: {n: t.clone().requires_grad_(True) for n, t in d.items()} for m, d in ppo.items()}
    summed = {m: {n: torch.zeros_like(t) for n, t in d.items()} for m, d in ppo.items()}
    for w in range(2):
        samples = []
        for h, c in enumerate(cpu[w]):
            c["returns"], c["value_preds"] = R.compute_returns(c["rewards"], c["value_preds"], c["masks"],
                                                               nv[w, h].reshape(1, 1))
            samples.append(R.gather_minibatch(c, R.normalized_advantages(c["returns"], c["value_preds"]),
                                              list(idx[0, w, h])))
        ref_l = R.update_policy(samples[0], samples[1], params)
        np.testing.assert_allclose(L[w].numpy(), np.array(ref_l), rtol=1e-3)
        for m in params:
            for n in params[m]:
                summed[m][n] += params[m][n].grad
    adam = {m: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
            for m, d in ppo.items()}
    pref = {m: {n: t.detach().clone() for n, t in d.items()} for m, d in ppo.items()}
    R.chief_step(pref, summed, adam, step=1)
    post = learner.state()
    a = torch.cat([post[m][n].flatten() for m in post for n in post[m]])
    b = torch.cat([pref[m][n].flatten() for m in post for n in post[m]])
    assert rel(a, b) < 2e-3          # theta after one step (TF32 gradients, Adam sign sensitivity)


def test_batched_actor_with_frame_cache_equals_single_env_act(agent):
    """E environments acted on in one batch, only the newest frame of each sliding 8-frame window
    encoded; features and values equal CadreAgent.act on the full windows (same kernels, row-independent math)."""
    from cadre_b200.actor import BatchedActor
    E, ticks_n = 3, 4
    rs = np.random.RandomState(7)
    streams = []
    for e in range(E):   # per env: 8 + ticks_n - 1 frames; tick t sees frames [t, t+8)
        n = 8 + ticks_n - 1
        streams.append(dict(rgb=rs.randint(0, 256, size=(n, 144, 256, 3)).astype(np.uint8),
                            route_fig=(rs.rand(n, 256, 144) < 0.1).astype(np.uint8) * 255,
                            measurements=rs.rand(n, 3)))
    actor = BatchedActor(agent, E)
    for t in range(ticks_n):
        ticks = [dict(rgb=s["rgb"][t:t + 8], route_fig=s["route_fig"][t:t + 8], measurements=s["measurements"][t:t + 8],
                      command=int((t + e) % 4)) for e, s in enumerate(streams)]
        before = actor.frames_encoded
        res = actor.act([{k: (v.copy() if hasattr(v, "copy") else v) for k, v in tk.items()} for tk in ticks])
        assert actor.frames_encoded - before == (8 * E if t == 0 else E)       # only the newest frames after tick 0
        for e in range(E):
            feat_b, act_b, lp_b, val_b, _ = res[e]
            tk = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in ticks[e].items()}
            feat_s, _, _, val_s, _ = agent.act(tk)
            assert torch.equal(feat_b, feat_s)
            for h in range(2):
                assert abs(val_b[h].item() - val_s[h].item()) < 1e-6
                assert 0 <= int(act_b[h]) < (33, 3)[h] and lp_b[h].shape == (1, 1)
    # a reset environment (its window does not continue the previous one) is re-encoded in full
    actor.reset(1)
    before = actor.frames_encoded
    actor.encode(ticks)
    assert actor.frames_encoded - before == 8 * E     # same ticks again: no window slid, everything re-encoded


@pytest.mark.parametrize("batched", [True, False])
def test_train_loop_runs_against_the_synthetic_env(batched):
    """cadre_b200.train.train (ppo_agent/train.py:53-110 + chief.py) for two short episodes with two logical workers:
    rollouts through BatchedActor or per-worker agent.act, GAE, 4 epochs x 2 minibatches of update_step."""
    from cadre_b200.config import load_config
    from cadre_b200.train import train
    cfg = load_config()
    cfg.rollout_cfg.num_steps = 8
    cfg.rollout_cfg.mini_batch_num = 2
    cfg.train_cfg.log_interval = 1000
    cfg.env_cfg["done_prob"] = 0.1          # exercise env resets / actor.reset inside the 16 ticks
    ppo0 = R.ppo_fixture_state(0)
    learner, hist = train(0, cfg.train_cfg, cfg.agent_cfg, cfg.env_cfg, cfg.rollout_cfg, R.danet_fixture_state(0),
                          ppo_state=ppo0, workers=2, max_episode=2, batched_acting=batched, log=lambda *a: None)
    assert len(hist) == 2 and all(np.isfinite(v) for ep in hist for v in ep)
    assert learner.step_count == 2 * cfg.train_cfg.ppo_epoch * 2
    new = learner.state()
    moved = (new["steer_lstm_0"]["rnn.weight_ih"].cpu() - ppo0["steer_lstm_0"]["rnn.weight_ih"]).abs().max().item()
    assert 0 < moved < 0.1                  # Adam moved the parameters by ~lr per step


def test_agent_from_unmodified_reference_config(tmp_path, monkeypatch):
    """`CadreAgent(**agent_cfg)` with the reference's config as it is (no extra keys, nothing injected): the
    perception checkpoint is found where models.py:54-63 + auto_danet.py:161-171 look for it
    ($CHALLENGE_DIR/carla_perception/Experiments34/.../net_epoch90), in the reference's file format, with off-path
    decoder tensors present."""
    from cadre_b200.agent import CadreAgent
    from cadre_b200.config import load_config
    from cadre_b200.models import default_pretrained_path
    monkeypatch.setenv("CHALLENGE_DIR", str(tmp_path))
    path = default_pretrained_path()
    os.makedirs(os.path.dirname(path))
    state = dict(R.danet_fixture_state(0))
    state["visual_branch.deconv.0.weight"] = torch.randn(4, 4, 3, 3)       # off the latent path
    torch.save({"epoch": 90, "metric": 0.0, "autoencoder": state}, path)
    cfg = load_config()
    assert "pretrained_path" not in cfg.agent_cfg.model_cfg
    torch.manual_seed(0)
    agent = CadreAgent(**cfg.agent_cfg)
    tick = R.synthetic_tick(np.random.RandomState(31))
    feat = agent.get_latent_feature(tick)
    with torch.no_grad():
        ref = R.agent_latent_feature(tick["rgb"], tick["route_fig"].copy(), tick["measurements"],
                                     R.danet_fixture_state(0))
    assert rel(feat, ref) < 1e-2
    assert agent.vae_params.networks["autoencoder"]["pretrained_path"] == path
    x = agent.pre_process(tick)
    assert np.array_equal(x.cpu().numpy(), R.pre_process(tick["rgb"], tick["route_fig"].copy()))


def test_snapshot_and_update_model_on_the_device(agent, tmp_path):
    """save_snapshot -> load_snapshot into a second agent -> identical act() values; update_model pulls parameters from
    a shared model list (agent.py:239-271)."""
    from cadre_b200.models import FlatParams, ModelDict
    path = str(tmp_path / "ppo_model_7.pt")
    agent.save_snapshot(path)
    other = ModelDict(FlatParams(agent.device, R.ppo_fixture_state(9)))
    keep = agent.owner.params.clone()
    try:
        agent.update_model(other)
        assert torch.equal(agent.owner.params, other.owner.params) and not torch.equal(agent.owner.params, keep)
        tick = R.synthetic_tick(np.random.RandomState(8))
        v_other = agent.get_value(False, (agent.get_latent_feature(tick), 1), (agent.get_latent_feature(tick), 2))
        agent.load_snapshot(path, agent.device)
        assert torch.equal(agent.owner.params, keep)
        v_back = agent.get_value(False, (agent.get_latent_feature(tick), 1), (agent.get_latent_feature(tick), 2))
        assert abs(v_other[0].item() - v_back[0].item()) > 1e-6         # different parameters, different value
    finally:
        agent.owner.params.copy_(keep)


def test_batched_actor_sampler_follows_the_policy_distribution(agent):
    """distributions.py:96-99 draws actions from softmax(logits); BatchedActor draws them for all environments with one
    multinomial per head. Over many ticks on the same observation the empirical action frequencies must match the
    policy's probabilities (chi-square, steer head folded to the bins with expected count >= 5)."""
    from cadre_b200.actor import BatchedActor
    E, ticks = 64, 40
    tick = R.synthetic_tick(np.random.RandomState(12))
    actor = BatchedActor(agent, E, verify_window=False)
    torch.manual_seed(123)
    feats, actions, logp, values = actor.act_batch([dict(tick, command=2) for _ in range(E)])
    # the policy's own probabilities for this observation / command, from the oracle
    ppo = R.ppo_fixture_state(0)
    counts = [torch.zeros(33), torch.zeros(3)]
    for _ in range(ticks):
        for e in range(E):
            actor.reset(e)                       # same 8-frame window every tick: the observation stays fixed
        _, actions, logp, _ = actor.act_batch([dict(tick, command=2) for _ in range(E)])
        for h in range(2):
            counts[h] += torch.bincount(actions[:, h].cpu(), minlength=counts[h].numel()).float()
    with torch.no_grad():
        for h, head in enumerate(("steer", "throttle")):
            f, _ = R.lstm_forward(feats[0].cpu(), torch.zeros(1, 530), torch.zeros(1, 530), ppo[f"{head}_lstm_2"])
            logits = R._mlp3(f, ppo[f"{head}_ppo_2"], "control.linear.")
            p = torch.softmax(logits, -1)[0]
            n = E * ticks
            exp = p * n
            big = exp >= 5
            chi = (((counts[h][big] - exp[big]) ** 2) / exp[big]).sum().item()
            if (~big).any():
                e_small, c_small = exp[~big].sum().item(), counts[h][~big].sum().item()
                chi += (c_small - e_small) ** 2 / max(e_small, 1e-9)
            dof = int(big.sum().item())
            assert chi < dof + 6 * (2 * dof) ** 0.5 + 10, (head, chi, dof)   # far beyond the 1e-6 tail of chi^2(dof)
            # log-prob returned with the action is the policy's
            a = int(actions[0, h])
            assert abs(logp[0, h].item() - torch.log(p[a]).item()) < 2e-2


def test_learn_staged_steps_equal_step_by_step_updates():
    """Learner.learn uploads the indices of all its update steps once and runs them from the device-side table (and
    the device-side Adam step counter); the parameters after three learn() phases (12 update steps, Adam's
    bias-corrected step sizes included) are bit-identical to driving update_step with the same indices one by one."""
    from cadre_b200.learner import Learner, RolloutPool
    ppo = R.ppo_fixture_state(0)
    cfg = dict(num_steps=64, mini_batch_num=2, feature_dims=530, seq_length=8, use_gae=True, gamma=0.99, tau=0.95)
    results = []
    for staged in (True, False):
        learner = Learner(2, 32, ppo, "cuda:0", seeds=[7, 8])
        pool = RolloutPool(2, cfg, "cuda:0")
        g = torch.Generator(device="cuda").manual_seed(3)
        b = pool.batched
        b["obs"].copy_(torch.randn(b["obs"].shape, device="cuda", generator=g))
        b["rewards"].copy_(torch.rand(b["rewards"].shape, device="cuda", generator=g))
        b["masks"].copy_((torch.rand(b["masks"].shape, device="cuda", generator=g) > 0.05).float())
        b["command"].copy_(torch.randint(0, 4, b["command"].shape, device="cuda", generator=g, dtype=torch.int32))
        b["action_log_probs"].copy_(-0.5 - 2.5 * torch.rand(b["action_log_probs"].shape, device="cuda", generator=g))
        b["value_preds"].copy_(torch.randn(b["value_preds"].shape, device="cuda", generator=g))
        b["action"].copy_(torch.randint(0, 3, b["action"].shape, device="cuda", generator=g))
        pool.compute_returns(torch.zeros(2, 2))
        for _ in range(3):
            if staged:
                assert learner.learn(pool, 2) == 4
            else:
                for _ in range(2):
                    for idx in learner.sample_epoch_indices(pool.storages):
                        learner.update_step(pool.storages, idx)
        learner.engine.check()
        assert learner.step_count == 12
        results.append((learner.params.clone(), learner.exp_avg_sq.clone()))
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])
