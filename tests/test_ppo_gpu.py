"""PPO path parity on the B200 (through the C ABI): GAE, routed forward, losses, gradients, clip + Adam,
against reference-generated goldens and the CPU oracle. Tolerances (BASELINE north_star): GAE 1e-5 relative
fp32; losses 1e-3 relative; parameters after the step by relative L2 (TF32 tensor-core operands)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(got, ref):
    got, ref = got.double().flatten().cpu(), ref.double().flatten().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


# ------------------------------------------------------------------------------------------------ GAE
@pytest.mark.parametrize("i,T,seed", [(0, 200, 0), (1, 200, 1), (2, 800, 2), (3, 7, 3)])
def test_gae_matches_reference_golden(golden_dir, i, T, seed):
    from cadre_b200 import ppo
    g = np.load(os.path.join(golden_dir, "gae.npz"))
    st = R.synthetic_storage(np.random.RandomState(seed), T=T, feature_dims=8, seq=1)
    if i == 1:
        st["masks"][::5] = 0.0
    r, v, m = (st[k].view(1, -1).to(DEV).contiguous() for k in ("rewards", "value_preds", "masks"))
    ret, adv = torch.zeros_like(r), torch.zeros(1, T, device=DEV)
    ppo.gae(r, v, m, torch.tensor([0.37 * (i + 1)], device=DEV), ret, adv)
    ref_ret = torch.from_numpy(g[f"returns_{i}"])[:T, 0]
    np.testing.assert_allclose(ret[0, :T].cpu().numpy(), ref_ret.numpy(), rtol=1e-5, atol=1e-5)
    assert rel(adv[0], torch.from_numpy(g[f"adv_{i}"])[:, 0]) < 1e-5
    assert v[0, T].item() == pytest.approx(0.37 * (i + 1))      # value_preds[-1] = next_value (storage.py:70)


@pytest.mark.parametrize("T", [2, 31, 32, 33, 255, 256, 257, 800, 1023, 1024, 1025, 3000])
def test_gae_all_kernel_variants_match_oracle(T):
    """Sequence lengths on both sides of every kernel switch (register-resident <= 256, <= 1024, general loop)
    and of the 32-lane chunking, with ~10 % episode boundaries; with and without advantage normalisation."""
    from cadre_b200 import ppo
    E = 5
    gen = torch.Generator().manual_seed(T)
    r = torch.rand(E, T + 1, generator=gen)
    v = torch.randn(E, T + 1, generator=gen)
    m = (torch.rand(E, T + 1, generator=gen) > 0.1).float()
    nv = torch.randn(E, generator=gen)
    for normalize in (True, False):
        vd = v.to(DEV).clone()
        ret, adv = torch.zeros(E, T + 1, device=DEV), torch.zeros(E, T, device=DEV)
        ppo.gae(r.to(DEV), vd, m.to(DEV), nv.to(DEV), ret, adv, normalize=normalize)
        for e in range(E):
            ref_ret, vp = R.compute_returns(r[e].view(-1, 1), v[e].view(-1, 1).clone(), m[e].view(-1, 1), nv[e].view(1, 1))
            np.testing.assert_allclose(ret[e, :T].cpu().numpy(), ref_ret[:T, 0].numpy(), rtol=1e-5, atol=1e-5)
            ref_adv = R.normalized_advantages(ref_ret, vp) if normalize else (ref_ret[:-1] - vp[:-1])
            assert rel(adv[e], ref_adv[:, 0]) < 2e-5
            assert vd[e, T].item() == pytest.approx(nv[e].item())


def test_gae_large_sweep_properties():
    """65 536 sequences x 1 024 steps (1.3 GB): linearity in the rewards and the all-masked closed form."""
    from cadre_b200 import ppo
    E, T = 65536, 1024
    g = torch.Generator(device=DEV).manual_seed(0)
    r = torch.rand(E, T + 1, device=DEV, generator=g)
    v = torch.randn(E, T + 1, device=DEV, generator=g)
    m = (torch.rand(E, T + 1, device=DEV, generator=g) > 0.02).float()
    nv = torch.randn(E, device=DEV, generator=g)
    out = [torch.zeros(E, T + 1, device=DEV) for _ in range(3)]
    adv = torch.zeros(E, T, device=DEV)
    ppo.gae(r, v.clone(), m, nv, out[0], adv, normalize=False)
    ppo.gae(2 * r, (2 * v).clone(), m, 2 * nv, out[1], adv, normalize=False)
    assert rel(out[1][:, :T], 2 * out[0][:, :T]) < 1e-6                      # linear in (r, V)
    ppo.gae(r, v.clone(), torch.zeros_like(m), nv, out[2], adv, normalize=False)
    assert torch.allclose(out[2][:, :T], r[:, :T], atol=1e-6)                # mask 0: returns_t = r_t
    # spot check 8 random sequences against the oracle loop
    for e in (0, 1, 777, 65535):
        ret, _ = R.compute_returns(r[e].cpu().view(-1, 1), v[e].cpu().view(-1, 1), m[e].cpu().view(-1, 1),
                                   nv[e].cpu().view(1, 1))
        np.testing.assert_allclose(out[0][e, :T].cpu().numpy(), ret[:T, 0].numpy(), rtol=2e-5, atol=2e-5)


# ------------------------------------------------------------------------------------------------ update
def _worker(w):
    rs = np.random.RandomState(100 + w)
    st_s = R.synthetic_storage(rs, actions=R.STEER_ACTIONS)
    st_t = R.synthetic_storage(rs, actions=R.THROTTLE_ACTIONS)
    if w == 1:
        for st in (st_s, st_t):
            st["hn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
            st["cn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
    advs = []
    for st, nv in ((st_s, 0.1), (st_t, -0.2)):
        st["returns"], st["value_preds"] = R.compute_returns(st["rewards"], st["value_preds"], st["masks"],
                                                             torch.tensor([[nv]]))
        advs.append(R.normalized_advantages(st["returns"], st["value_preds"]))
    torch.manual_seed(500 + w)
    return (st_s, st_t), advs, (R.minibatch_indices()[0], R.minibatch_indices()[0])


def _dev(st):
    return SimpleNamespace(**{k: v.to(DEV).contiguous() for k, v in st.items()})


@pytest.fixture(scope="module")
def two_worker_update():
    from cadre_b200 import ppo, ppo_params
    torch.set_num_threads(8)
    sd = R.ppo_fixture_state(0)
    cpu = [_worker(w) for w in range(2)]
    storages = [(_dev(c[0][0]), _dev(c[0][1])) for c in cpu]
    advs = [(c[1][0].to(DEV).contiguous(), c[1][1].to(DEV).contiguous()) for c in cpu]
    idx = np.array([[c[2][0], c[2][1]] for c in cpu], dtype=np.int32)
    flat = ppo_params.pack_state(sd, DEV)
    grads = torch.zeros_like(flat)
    eng = ppo.PpoEngine(2, 100, device=DEV)
    ev = eng.evaluate(storages, advs, idx, flat).cpu()
    losses = eng.update(storages, advs, idx, flat, grads).cpu()
    # oracle: per-worker update_policy, gradients summed like Shared_grad_buffers.add_gradient
    params = {m: {n: t.clone().requires_grad_(True) for n, t in d.items()} for m, d in sd.items()}
    summed = {m: {n: torch.zeros_like(t) for n, t in d.items()} for m, d in sd.items()}
    ref_losses = []
    for w in range(2):
        s = R.gather_minibatch(cpu[w][0][0], cpu[w][1][0], cpu[w][2][0])
        t = R.gather_minibatch(cpu[w][0][1], cpu[w][1][1], cpu[w][2][1])
        ref_losses.append(R.update_policy(s, t, params))
        for m in params:
            for n in params[m]:
                summed[m][n] += params[m][n].grad
    return dict(eng=eng, flat=flat, grads=grads, losses=losses, ev=ev, cpu=cpu, sd=sd, summed=summed,
                ref_losses=ref_losses, ppo_params=ppo_params)


def test_losses_match_reference_golden(two_worker_update, golden_dir):
    u = two_worker_update
    g = np.load(os.path.join(golden_dir, "update.npz"))
    for w in range(2):
        L = u["losses"][w].sum(0)
        got = np.array([0.1 * L[0].item(), L[1].item(), 0.01 * L[2].item()])
        np.testing.assert_allclose(got, g["losses"][w], rtol=1e-3)                 # reference run (golden)
        np.testing.assert_allclose(got, np.array(u["ref_losses"][w]), rtol=1e-3)   # oracle


def test_forward_rows_match_oracle(two_worker_update):
    u = two_worker_update
    sd = u["sd"]
    for head_i, head in enumerate(("steer", "throttle")):
        ref = []
        for w in range(2):
            obs, action, _, _, _, _, _, (hn, cn), command = R.gather_minibatch(u["cpu"][w][0][head_i],
                                                                               u["cpu"][w][1][head_i],
                                                                               u["cpu"][w][2][head_i])
            acc = torch.zeros(100, 3)
            with torch.no_grad():
                for c in range(4):
                    feat, _ = R.lstm_forward(obs.clone(), hn, cn, sd[f"{head}_lstm_{c}"])
                    v, lp, en = R.evaluate_actions(feat, action, sd[f"{head}_ppo_{c}"])
                    acc += torch.cat([v, lp, en], 1) * (command == c)
            ref.append(acc)
        ref = torch.cat(ref, 0)
        got = u["ev"][head_i, :, :3]
        assert rel(got[:, 0], ref[:, 0]) < 2e-3        # value (TF32 operands)
        assert rel(got[:, 1], ref[:, 1]) < 1e-3        # log-prob of the stored action
        assert rel(got[:, 2], ref[:, 2]) < 1e-4        # entropy


def test_gradients_match_oracle_and_golden(two_worker_update, golden_dir):
    u = two_worker_update
    P = u["ppo_params"]
    g = P.unpack_state(u["grads"].cpu())
    names = [(m, n) for m in P.MODULE_ORDER for n in P.module_param_names(m)]
    got = torch.cat([g[m][n].flatten() for m, n in names])
    ref = torch.cat([u["summed"][m][n].flatten() for m, n in names])
    assert rel(got, ref) < 2e-2                                                   # all 19.4 M gradients
    per = sorted(rel(g[m][n], u["summed"][m][n]) for m, n in names)
    assert per[len(per) // 2] < 5e-3 and per[-1] < 1e-1
    # layout padding never receives gradient
    probe = torch.zeros(P.TOTAL)
    for m, n in names:
        P.tensor_view(probe, m, n)[0].fill_(1.0)
    assert u["grads"].cpu()[probe == 0].abs().max().item() == 0.0
    # unused experts get exactly zero gradient (reference: masked rows contribute 0)
    gold = np.load(os.path.join(golden_dir, "update.npz"))
    assert gold["w0_grad_stats"].shape[0] == 128


def test_clip_adam_matches_oracle(two_worker_update):
    u = two_worker_update
    P = u["ppo_params"]
    flat, grads = u["flat"].clone(), u["grads"]
    m1, m2 = torch.zeros_like(flat), torch.zeros_like(flat)
    # feed the ORACLE's summed gradient so that this test isolates the clip + Adam kernels
    gflat = P.pack_state(u["summed"], DEV)
    u["eng"].adam_step(flat, gflat, m1, m2, step=1)
    sd = u["sd"]
    adam = {m: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
            for m, d in sd.items()}
    pref = {m: {n: t.detach().clone() for n, t in d.items()} for m, d in sd.items()}
    R.chief_step(pref, u["summed"], adam, step=1)
    post = P.unpack_state(flat.cpu())
    names = [(m, n) for m in P.MODULE_ORDER for n in P.module_param_names(m)]
    a = torch.cat([post[m][n].flatten() for m, n in names])
    b = torch.cat([pref[m][n].flatten() for m, n in names])
    a0 = torch.cat([sd[m][n].flatten() for m, n in names])
    assert rel(a, b) < 1e-6 and rel(a - a0, b - a0) < 1e-3
    norms = u["eng"].module_norms()
    for m in sd:
        ref_n = torch.sqrt(sum((u["summed"][m][n].double() ** 2).sum() for n in sd[m])).item()
        assert norms[m] == pytest.approx(ref_n, rel=1e-5)
    # a second step with a huge gradient exercises the clip branch (coef < 1) per module
    big = gflat * 1e5
    u["eng"].adam_step(flat, big, m1, m2, step=2)
    grads2 = {m: {n: u["summed"][m][n] * 1e5 for n in sd[m]} for m in sd}
    R.chief_step(pref, grads2, adam, step=2)
    post = P.unpack_state(flat.cpu())
    a = torch.cat([post[m][n].flatten() for m, n in names])
    b = torch.cat([pref[m][n].flatten() for m, n in names])
    assert rel(a, b) < 1e-6


def test_end_to_end_step_parameters(two_worker_update, golden_dir):
    """CUDA gradients -> CUDA clip + Adam vs the reference's post-step parameters (golden)."""
    u = two_worker_update
    P = u["ppo_params"]
    g = np.load(os.path.join(golden_dir, "update.npz"))
    flat = u["flat"].clone()
    m1, m2 = torch.zeros_like(flat), torch.zeros_like(flat)
    u["eng"].adam_step(flat, u["grads"], m1, m2, step=1)
    post = P.unpack_state(flat.cpu())
    k = 0
    num = den = 0.0
    for m in P.MODULE_ORDER:
        for n in P.module_param_names(m):
            sl = post[m][n].flatten()[:32].double()
            ref = torch.from_numpy(g["post_param_slices"][k][:len(sl)]).double()
            num += ((sl - ref) ** 2).sum().item()
            den += (ref ** 2).sum().item()
            # the first Adam step moves every weight by ~lr * sign(g): bounded by 2 * lr elementwise
            assert (sl - ref).abs().max().item() <= 2 * 3e-4 + 1e-6
            k += 1
    assert (num / den) ** 0.5 < 5e-3
