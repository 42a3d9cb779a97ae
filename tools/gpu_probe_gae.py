"""GAE roofline sweep only (20 algorithmic bytes per (sequence, step))."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import ppo
dev = "cuda:0"
for (E, T) in ((65536, 1024), (262144, 256), (16384, 4096), (65536, 1000)):
    g = torch.Generator(device=dev).manual_seed(0)
    r = torch.rand(E, T + 1, device=dev, generator=g); v = torch.randn(E, T + 1, device=dev, generator=g)
    m = (torch.rand(E, T + 1, device=dev, generator=g) > 0.02).float(); nv = torch.randn(E, device=dev, generator=g)
    ret = torch.zeros(E, T + 1, device=dev); adv = torch.zeros(E, T, device=dev)
    for _ in range(3): ppo.gae(r, v, m, nv, ret, adv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): ppo.gae(r, v, m, nv, ret, adv)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = 20.0 * E * T / (ms * 1e-3) / 1e9
    print(f"gae E={E} T={T}: {ms:.3f} ms  {gbs:.0f} GB/s  {gbs/6453.1:.3f} of measured HBM peak", flush=True)
    del r, v, m, ret, adv
