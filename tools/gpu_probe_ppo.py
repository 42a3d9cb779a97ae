"""GPU probe of the PPO path: GAE, forward (evaluate), losses, gradients and clip+Adam vs the CPU oracle.
Prints diagnostics; the pytest versions live in tests/test_ppo_gpu.py."""
import os
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import ppo, ppo_params  # noqa: E402
from oracle import restate as R  # noqa: E402

dev = torch.device("cuda:0")


def rel(got, ref):
    got, ref = got.double().flatten().cpu(), ref.double().flatten().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-30)).item()


def worker_storages(w, nonzero_state=False):
    rs = np.random.RandomState(100 + w)
    st_s = R.synthetic_storage(rs, actions=R.STEER_ACTIONS)
    st_t = R.synthetic_storage(rs, actions=R.THROTTLE_ACTIONS)
    if w == 1 or nonzero_state:
        for st in (st_s, st_t):
            st["hn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
            st["cn"] = torch.from_numpy(rs.randn(201, 530).astype(np.float32) * 0.3)
    advs = []
    for st, nv in ((st_s, 0.1), (st_t, -0.2)):
        st["returns"], st["value_preds"] = R.compute_returns(st["rewards"], st["value_preds"], st["masks"],
                                                             torch.tensor([[nv]]))
        advs.append(R.normalized_advantages(st["returns"], st["value_preds"]))
    torch.manual_seed(500 + w)
    idx_s = R.minibatch_indices()[0]
    idx_t = R.minibatch_indices()[0]
    return (st_s, st_t), advs, (idx_s, idx_t)


def to_dev(st):
    return SimpleNamespace(**{k: v.to(dev).contiguous() for k, v in st.items()})


def main():
    torch.set_num_threads(8)
    # ---- GAE
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "gae.npz"))
    for i, (T, seed) in enumerate(((200, 0), (200, 1), (800, 2), (7, 3))):
        st = R.synthetic_storage(np.random.RandomState(seed), T=T, feature_dims=8, seq=1)
        if i == 1:
            st["masks"][::5] = 0.0
        r = st["rewards"].view(1, -1).to(dev).contiguous()
        v = st["value_preds"].view(1, -1).to(dev).contiguous()
        m = st["masks"].view(1, -1).to(dev).contiguous()
        ret = torch.zeros_like(r)
        adv = torch.zeros(1, T, device=dev)
        ppo.gae(r, v, m, torch.tensor([0.37 * (i + 1)], device=dev), ret, adv)
        print(f"gae[{i}] T={T}: returns rel {rel(ret[0, :T], torch.from_numpy(g[f'returns_{i}'])[:T, 0]):.2e} "
              f"adv rel {rel(adv[0], torch.from_numpy(g[f'adv_{i}'])[:, 0]):.2e}", flush=True)

    # ---- update, W = 2
    sd = R.ppo_fixture_state(0)
    flat = ppo_params.pack_state(sd, dev)
    W, mb = 2, 100
    eng = ppo.PpoEngine(W, mb, device=dev)
    cpu = [worker_storages(w) for w in range(W)]
    storages = [(to_dev(c[0][0]), to_dev(c[0][1])) for c in cpu]
    advs = [(c[1][0].to(dev).contiguous(), c[1][1].to(dev).contiguous()) for c in cpu]
    idx = np.array([[c[2][0], c[2][1]] for c in cpu], dtype=np.int32)

    # oracle
    params = {m_: {n: t.clone().requires_grad_(True) for n, t in d.items()} for m_, d in sd.items()}
    summed = {m_: {n: torch.zeros_like(t) for n, t in d.items()} for m_, d in sd.items()}
    ref_losses, ref_rows = [], []
    for w in range(W):
        s_samp = R.gather_minibatch(cpu[w][0][0], cpu[w][1][0], cpu[w][2][0])
        t_samp = R.gather_minibatch(cpu[w][0][1], cpu[w][1][1], cpu[w][2][1])
        ref_losses.append(R.update_policy(s_samp, t_samp, params))
        for m_ in params:
            for n in params[m_]:
                summed[m_][n] += params[m_][n].grad
        # per-row forward reference (routed == dense-masked)
        with torch.no_grad():
            for head, samp in (("steer", s_samp), ("throttle", t_samp)):
                obs, action, _, _, _, _, _, (hn, cn), command = samp
                v_all = torch.zeros(mb, 1)
                lp_all = torch.zeros(mb, 1)
                en_all = torch.zeros(mb, 1)
                for c in range(4):
                    feat, _ = R.lstm_forward(obs.clone(), hn, cn, sd[f"{head}_lstm_{c}"])
                    vv, lp, en = R.evaluate_actions(feat, action, sd[f"{head}_ppo_{c}"])
                    sel = command == c
                    v_all += vv * sel
                    lp_all += lp * sel
                    en_all += en * sel
                ref_rows.append(torch.cat([v_all, lp_all, en_all], 1))
    ev = eng.evaluate(storages, advs, idx, flat).cpu()
    # row order of the library: [head][worker*mb + i]
    for head in range(2):
        ref = torch.cat([ref_rows[w * 2 + head] for w in range(W)], 0)
        got = ev[head, :, :3]
        print(f"evaluate head {head}: value rel {rel(got[:, 0], ref[:, 0]):.2e} logp rel {rel(got[:, 1], ref[:, 1]):.2e} "
              f"entropy rel {rel(got[:, 2], ref[:, 2]):.2e}", flush=True)

    grads = torch.zeros_like(flat)
    t0 = time.time()
    losses = eng.update(storages, advs, idx, flat, grads)
    torch.cuda.synchronize()
    print("update wall (first call) ms", (time.time() - t0) * 1e3, "launches", eng.launches)
    L = losses.cpu()
    for w in range(W):
        got = [0.1 * (L[w, 0, 0] + L[w, 1, 0]).item(), (L[w, 0, 1] + L[w, 1, 1]).item(),
               0.01 * (L[w, 0, 2] + L[w, 1, 2]).item()]
        print(f"losses w{w}: got {got} ref {ref_losses[w]} rel "
              f"{[abs(a - b) / abs(b) for a, b in zip(got, ref_losses[w])]}", flush=True)
    gstate = ppo_params.unpack_state(grads.cpu())
    worst = []
    for m_ in ppo_params.MODULE_ORDER:
        for n in ppo_params.module_param_names(m_):
            worst.append((rel(gstate[m_][n], summed[m_][n]), m_, n, summed[m_][n].norm().item()))
    worst.sort(reverse=True)
    print("grad rel-L2 worst 8:", [(f"{a:.2e}", b, c, f"{d:.2e}") for a, b, c, d in worst[:8]])
    print("grad rel-L2 median:", f"{sorted(x[0] for x in worst)[len(worst) // 2]:.2e}")
    allg = torch.cat([gstate[m_][n].flatten() for m_ in ppo_params.MODULE_ORDER
                      for n in ppo_params.module_param_names(m_)])
    allr = torch.cat([summed[m_][n].flatten() for m_ in ppo_params.MODULE_ORDER
                      for n in ppo_params.module_param_names(m_)])
    print("grad rel-L2 global:", f"{rel(allg, allr):.2e}", flush=True)
    # padding must stay zero
    mask = torch.ones(ppo_params.TOTAL, dtype=torch.bool)
    probe = torch.zeros(ppo_params.TOTAL)
    for m_ in ppo_params.MODULE_ORDER:
        for n in ppo_params.module_param_names(m_):
            v_, _ = ppo_params.tensor_view(probe, m_, n)
            v_.fill_(1.0)
    pad = probe == 0
    print("pad elements", int(pad.sum()), "max |grad| on padding", grads.cpu()[pad].abs().max().item())

    # ---- clip + Adam
    m1 = torch.zeros_like(flat)
    m2 = torch.zeros_like(flat)
    before = flat.clone()
    eng.adam_step(flat, grads, m1, m2, step=1)
    torch.cuda.synchronize()
    adam = {m_: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
            for m_, d in sd.items()}
    pref = {m_: {n: t.detach().clone() for n, t in d.items()} for m_, d in sd.items()}
    R.chief_step(pref, summed, adam, step=1)
    post = ppo_params.unpack_state(flat.cpu())
    pre = ppo_params.unpack_state(before.cpu())
    a = torch.cat([post[m_][n].flatten() for m_ in post for n in post[m_]])
    b = torch.cat([pref[m_][n].flatten() for m_ in post for n in post[m_]])
    a0 = torch.cat([pre[m_][n].flatten() for m_ in post for n in post[m_]])
    print(f"adam: theta rel {rel(a, b):.2e}  delta rel {rel(a - a0, b - a0):.2e}  max|dtheta| {(a - a0).abs().max():.2e}")
    norms = eng.module_norms()
    ref_norm = {m_: torch.sqrt(sum((summed[m_][n].double() ** 2).sum() for n in summed[m_])).item() for m_ in summed}
    print("module norm rel:", max(abs(norms[m_] - ref_norm[m_]) / ref_norm[m_] for m_ in norms))

    # ---- timing (kernel only)
    for _ in range(3):
        eng.update(storages, advs, idx, flat, grads)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        eng.update(storages, advs, idx, flat, grads)
        eng.adam_step(flat, grads, m1, m2, step=2)
    e1.record()
    torch.cuda.synchronize()
    print(f"update+adam W={W} mb={mb}: {e0.elapsed_time(e1) / 10:.3f} ms per step")


if __name__ == "__main__":
    main()
