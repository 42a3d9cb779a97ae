"""Encode phase of the end-to-end step (RolloutIngest.encode(unique=True) from pinned host frames, cfg 3: 4 workers x 207
distinct frames) for different chunk ramps (CADRE_INGEST_RAMP) and chunk sizes; device-resident encode for comparison."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import fixtures as FX
from cadre_b200.ingest import RolloutIngest, chunk_schedule
from cadre_b200.learner import RolloutPool
dev = "cuda:0"
W, T, S = 4, 200, 8
K = T + S - 1
g = torch.Generator().manual_seed(0)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
rgb = pin(torch.randint(0, 256, (W, K, 144, 256, 3), dtype=torch.uint8, generator=g))
route = pin((torch.rand(W, K, 256, 144, generator=g) < 0.1).to(torch.uint8) * 255)
meas = pin(torch.rand(W, K, 3, dtype=torch.float64, generator=g))
pool = RolloutPool(W, dict(num_steps=T, mini_batch_num=2, feature_dims=530, seq_length=S, use_gae=True, gamma=0.99, tau=0.95), dev)
obs = pool.batched["obs"]
out = {}
def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
sd = FX.danet_fixture_state(0)
for chunk in (640, 448, 320):
    ing = RolloutIngest(sd, dev, W, T, S, 530, max_chunk=chunk, streams=2)
    for ramp in ("128,512", "128", "64,256", "64,192", "96,224", "32,96,288", "64,128,256", "256", ""):
        os.environ["CADRE_INGEST_RAMP"] = ramp
        ms = timed(lambda: ing.encode(rgb, route, meas, obs, unique=True))
        out[f"chunk{chunk}_ramp[{ramp}]"] = {"ms": round(ms, 4), "schedule": chunk_schedule(W * K, chunk)}
        print(f"chunk {chunk} ramp [{ramp}] -> {chunk_schedule(W * K, chunk)}: {ms:.3f} ms", flush=True)
    d = (rgb.cuda(), route.cuda(), meas.cuda())
    ms = timed(lambda: ing.encode(d[0], d[1], d[2], obs, unique=True))
    out[f"chunk{chunk}_device_resident"] = round(ms, 4)
    print(f"chunk {chunk} device-resident: {ms:.3f} ms", flush=True)
    del ing, d
json.dump(out, open("gpurun_out/r2_ingest_ramp.json", "w"), indent=1)
