"""Extra measurements for profiles/: GAE roofline sweep, clip+Adam roofline, encoder batch sweep (config 2),
scaled PPO update (config 5 per-GPU share: 8 workers x mb=400 -> 3200 rows per head)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import ppo, ppo_params
from cadre_b200.encoder import Encoder
from cadre_b200.learner import Learner, RolloutPool
from cadre_b200 import fixtures as R
dev = "cuda:0"
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {"hbm_gbs": 6453.1, "bf16_tflops_sustained": 1417.3}
out = {}

def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

# ---- GAE sweep: 20 B per (sequence, step)
E, T = 65536, 1024
g = torch.Generator(device=dev).manual_seed(0)
r = torch.rand(E, T + 1, device=dev, generator=g); v = torch.randn(E, T + 1, device=dev, generator=g)
m = (torch.rand(E, T + 1, device=dev, generator=g) > 0.02).float(); nv = torch.randn(E, device=dev, generator=g)
ret = torch.zeros(E, T + 1, device=dev); adv = torch.zeros(E, T, device=dev)
ms = timed(lambda: ppo.gae(r, v, m, nv, ret, adv))
gbs = 20.0 * E * T / (ms * 1e-3) / 1e9
out["gae_sweep"] = {"sequences": E, "steps": T, "ms": ms, "GBps": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]}
print("gae sweep", out["gae_sweep"], flush=True)
for (E2, T2) in ((8, 200), (128, 800)):
    r2 = torch.rand(E2, T2 + 1, device=dev); v2 = torch.randn(E2, T2 + 1, device=dev); m2 = torch.ones(E2, T2 + 1, device=dev)
    ms2 = timed(lambda: ppo.gae(r2, v2, m2, torch.zeros(E2, device=dev), torch.zeros(E2, T2 + 1, device=dev), torch.zeros(E2, T2, device=dev)), iters=50)
    out[f"gae_cfg_E{E2}_T{T2}_us"] = ms2 * 1e3
    print(f"gae E={E2} T={T2}: {ms2*1e3:.1f} us (latency regime)", flush=True)
del r, v, m, ret, adv

# ---- clip + Adam: 32 B / parameter
flat = torch.randn(ppo_params.TOTAL, device=dev) * 0.01
grads = torch.randn_like(flat) * 1e-3; m1 = torch.zeros_like(flat); m2 = torch.zeros_like(flat)
eng = ppo.PpoEngine(1, 8, device=dev)
ms = timed(lambda: eng.adam_step(flat, grads, m1, m2, step=3), iters=20)
gbs = 32.0 * ppo_params.TOTAL / (ms * 1e-3) / 1e9
out["clip_adam"] = {"params": ppo_params.TOTAL, "ms": ms, "GBps": gbs, "frac_of_measured_hbm": gbs / peaks["hbm_gbs"]}
print("clip+adam", out["clip_adam"], flush=True)

# ---- encoder batch sweep (BASELINE config 2)
enc = Encoder(R.danet_fixture_state(0), dev, max_batch=1024)
sweep = []
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
    x = torch.rand(B, 4, 144, 256, device=dev); o = torch.empty(B, 512, device=dev)
    ms = timed(lambda: enc.forward_f32(x, o), iters=10)
    sweep.append({"batch": B, "ms": round(ms, 4), "frames_per_s": round(B / ms * 1e3), "tflops_equiv_3.0875G": round(3.0875 * B / ms, 1)})
    print(sweep[-1], flush=True)
out["encoder_sweep"] = sweep
del enc

# ---- scaled PPO update (config 5 per-GPU share)
for (W, Tn, mbn) in ((4, 200, 2), (8, 800, 2)):
    mb = Tn // mbn
    learner = Learner(W, mb, R.ppo_fixture_state(0), dev, seeds=list(range(W)))
    pool = RolloutPool(W, dict(num_steps=Tn, mini_batch_num=mbn, feature_dims=530, seq_length=8, use_gae=True, gamma=0.99, tau=0.95), dev)
    b = pool.batched
    gg = torch.Generator(device=dev).manual_seed(1)
    b["obs"].copy_(torch.randn(b["obs"].shape, device=dev, generator=gg)); b["rewards"].copy_(torch.rand(b["rewards"].shape, device=dev, generator=gg))
    b["masks"].fill_(1.0); b["command"].copy_(torch.randint(0, 4, b["command"].shape, device=dev, generator=gg, dtype=torch.int32))
    b["action_log_probs"].fill_(-1.5); b["value_preds"].copy_(torch.randn(b["value_preds"].shape, device=dev, generator=gg))
    for w in range(W):
        b["action"][2 * w].copy_(torch.randint(0, 33, (Tn + 1, 1), device=dev, generator=gg)); b["action"][2 * w + 1].copy_(torch.randint(0, 3, (Tn + 1, 1), device=dev, generator=gg))
    pool.compute_returns(torch.zeros(W, 2, device=dev))
    idx = learner.sample_epoch_indices(pool.storages)
    ms = timed(lambda: learner.update_step(pool.storages, idx[0]), iters=10)
    rows = W * mb
    fl = 217.8e6 * rows
    out[f"ppo_update_W{W}_mb{mb}"] = {"rows_per_head": rows, "ms_per_update_step": ms, "samples_per_s": rows / (ms * 1e-3),
                                      "tflops_routed_217.8MF_per_row": fl / (ms * 1e-3) / 1e12}
    print(f"ppo update W={W} mb={mb}", out[f"ppo_update_W{W}_mb{mb}"], flush=True)
    del learner, pool
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2_extras.json", "w"), indent=1)
