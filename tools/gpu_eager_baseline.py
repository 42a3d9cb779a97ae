"""GPU-side bar: the reference graph (oracle/restate.py, plain PyTorch modules' formulas) run in
torch eager on the B200 -- cuDNN / cuBLAS library path: fp32 with torch's default TF32 convolution setting, bf16
autocast, and bf16 autocast on channels_last tensors -- next to the same workload through cadre_b200.
A baseline tool, not product code. Default: a batch sweep written to gpurun_out/r2_eager_baseline.json;
`--json-line` (used by bench.py in a SUBPROCESS, so that no library kernel is loaded into the measured process): batch
640 + the PPO update step only, one JSON line on stdout."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import restate as R
JSON_LINE = "--json-line" in sys.argv
if not JSON_LINE:
    from cadre_b200.encoder import Encoder
dev = torch.device("cuda:0")
out = {"torch": torch.__version__, "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32,
       "matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32}


def timed(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


sd_cpu = R.danet_fixture_state(0)
sd = {k: v.to(dev) for k, v in sd_cpu.items()}
sd_cl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd.items()}
enc = None if JSON_LINE else Encoder(sd_cpu, "cuda:0", max_batch=640)
rows = []
for B in ((640,) if JSON_LINE else (8, 64, 256, 640)):
    x = torch.rand(B, 4, 144, 256, device=dev)
    x_cl = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        ms_ref = timed(lambda: R.encoder_latent(x, sd), 5)
        torch.backends.cudnn.benchmark = True
        ms_ref_b = timed(lambda: R.encoder_latent(x, sd), 5, warm=5)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_ref_bf16 = timed(lambda: R.encoder_latent(x, sd), 5, warm=5)
            ms_ref_bf16_cl = timed(lambda: R.encoder_latent(x_cl, sd_cl), 5, warm=5)
        torch.backends.cudnn.benchmark = False
    row = {"batch": B, "eager_fp32_ms": round(ms_ref, 3), "eager_fp32_cudnn_benchmark_ms": round(ms_ref_b, 3),
           "eager_autocast_bf16_ms": round(ms_ref_bf16, 3),
           "eager_autocast_bf16_channels_last_ms": round(ms_ref_bf16_cl, 3)}
    if enc is not None:
        o = torch.empty(B, 512, device=dev)
        ms_new = timed(lambda: enc.forward_f32(x, o), 10)
        row["cadre_b200_ms"] = round(ms_new, 3)
        row["speedup_vs_best_eager"] = round(min(ms_ref, ms_ref_b, ms_ref_bf16, ms_ref_bf16_cl) / ms_new, 2)
    rows.append(row)
    if not JSON_LINE:
        print(rows[-1], flush=True)
out["encoder"] = rows
del enc

# PPO update step: W=4 workers x (steer, throttle) minibatches of 100 rows, reference-dense update_policy + chief
psd = R.ppo_fixture_state(0)
params = {m: {n: t.clone().to(dev).requires_grad_(True) for n, t in d.items()} for m, d in psd.items()}
adam = {m: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
        for m, d in params.items()}
rs = np.random.RandomState(0)
W, mb = 4, 100
samples = []
for w in range(W):
    pair = []
    for a in (R.STEER_ACTIONS, R.THROTTLE_ACTIONS):
        st = R.synthetic_storage(rs, T=200, actions=a)
        st["returns"], vp = R.compute_returns(st["rewards"], st["value_preds"], st["masks"], torch.tensor([[0.1]]))
        adv = R.normalized_advantages(st["returns"], vp)
        mbt = R.gather_minibatch(st, adv, list(range(mb)))
        pair.append(tuple(t.to(dev) if torch.is_tensor(t) else [u.to(dev) for u in t] if isinstance(t, (list, tuple)) else t
                          for t in mbt))
    samples.append(pair)
step = [0]


def ref_update():
    summed = None
    for w in range(W):
        for d in params.values():
            for p in d.values(): p.grad = None
        R.update_policy(samples[w][0], samples[w][1], params)
        g = {m: {n: p.grad.clone() for n, p in d.items()} for m, d in params.items()}
        if summed is None: summed = g
        else:
            for m in g:
                for n in g[m]: summed[m][n] += g[m][n]
    step[0] += 1
    R.chief_step(params, summed, adam, step=step[0])


try:
    ms = timed(ref_update, 5, warm=2)
    out["ppo_update_eager_ms"] = round(ms, 3)
    if not JSON_LINE:
        print("eager PPO update step (4 workers, dense 4-command formulation)", ms, "ms", flush=True)
except Exception as e:  # the restatement is CPU-first; report rather than fail
    out["ppo_update_eager_error"] = repr(e)[:300]
    if not JSON_LINE:
        print("eager PPO update failed:", repr(e)[:300], flush=True)
if JSON_LINE:
    r = rows[0]
    print(json.dumps({"what": "reference graph in torch eager (cuDNN / cuBLAS) on the same GPU, separate process",
                      "torch": out["torch"], "encoder_b640_fp32_ms": r["eager_fp32_ms"],
                      "encoder_b640_bf16_ms": r["eager_autocast_bf16_ms"],
                      "encoder_b640_bf16_channels_last_ms": r["eager_autocast_bf16_channels_last_ms"],
                      "ppo_update_step_fp32_ms": out.get("ppo_update_eager_ms"),
                      "ppo_update_error": out.get("ppo_update_eager_error")}))
else:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r2_eager_baseline.json", "w"), indent=1)
