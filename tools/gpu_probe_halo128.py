"""cadre_conv3x3_flat128 (tc_halo128.cuh) against torch fp32 conv and the implicit-GEMM kernel: error breakdown per tile
position / channel half (debug aid) and device time of both kernels at the benchmarked chunk (B = 640)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
L = _lib.lib()
dt = _lib.enc_dtype()
dev = "cuda:0"
H, W, C = 18, 32, 128

def run(B, res):
    g = torch.Generator(device=dev).manual_seed(B)
    x = torch.randn(B, C, H, W, device=dev, generator=g).to(dt)
    w = (torch.randn(C, C, 3, 3, device=dev, generator=g) / (C * 9) ** 0.5).to(dt)
    bias = torch.randn(C, device=dev, generator=g)
    xp = torch.zeros(B, H + 2, W + 2, C, device=dev, dtype=dt)
    xp[:, 1:-1, 1:-1] = x.permute(0, 2, 3, 1)
    w_k = w.permute(0, 2, 3, 1).contiguous().view(C, -1)
    r = rp = None
    if res:
        r = torch.randn(B, H, W, C, device=dev, generator=g).to(dt)
        rp = torch.zeros_like(xp)
        rp[:, 1:-1, 1:-1] = r
    out = torch.full((B, H + 2, W + 2, C), 7.0, device=dev, dtype=dt)
    rc = L.cadre_conv3x3_flat128(_lib.ptr(xp), B, H, W, _lib.ptr(w_k), _lib.ptr(bias), _lib.ptr(rp), 1, _lib.ptr(out),
                                 _lib.stream_ptr())
    if rc:
        print("rc", rc, L.cadre_last_error().decode()); return
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), bias, stride=1, padding=1)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    ref = ref.relu().permute(0, 2, 3, 1)
    refp = torch.zeros(B, H + 2, W + 2, C, device=dev)
    refp[:, 1:-1, 1:-1] = ref
    d = (out.float() - refp).view(-1, C)
    rel = (d.norm() / refp.norm()).item()
    print(f"B={B} res={res}: rel_l2 {rel:.3e}  max|d| {d.abs().max().item():.3e}  max|ref| {refp.abs().max().item():.3f}"
          f"  untouched(7.0) {(out == 7.0).sum().item()}  nan {torch.isnan(out.float()).sum().item()}", flush=True)
    if rel > 1e-3:
        P = d.shape[0]
        pad = (-P) % 256
        dd = torch.cat([d, torch.zeros(pad, C, device=dev)]).view(-1, 2, 128, 2, 64)    # [group][tile s][row][half][ch]
        e = dd.pow(2).sum(dim=(2, 4)).sqrt()                                             # [group][s][half]
        print("  error by (tile of the group, channel half):", e.sum(0).tolist())
        er = dd.pow(2).sum(dim=(0, 1, 3, 4)).sqrt()
        print("  error by row within the tile (first 16 / last 16):", [round(v, 3) for v in er[:16].tolist()],
              [round(v, 3) for v in er[-16:].tolist()])
        eg = dd.pow(2).sum(dim=(1, 2, 3, 4)).sqrt()
        bad = (eg > 1e-2).nonzero().flatten()[:20].tolist()
        print("  first bad groups:", bad, "of", eg.numel())

def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3

for B, res in ((1, False), (3, True), (37, True), (640, False), (640, True)):
    run(B, res)
B = 640
xp = torch.randn(B, H + 2, W + 2, C, device=dev).to(dt)
rp = torch.randn(B, H + 2, W + 2, C, device=dev).to(dt)
w_k = (torch.randn(C, 9 * C, device=dev) / 34.0).to(dt)
bias = torch.randn(C, device=dev)
out = torch.empty(B, H + 2, W + 2, C, device=dev, dtype=dt)
old = torch.empty(B, H, W, C, device=dev, dtype=dt)
r_old = torch.randn(B, H, W, C, device=dev).to(dt)
fl = 2.0 * B * H * W * C * C * 9
for res in (False, True):
    us = timed(lambda: L.cadre_conv3x3_flat128(_lib.ptr(xp), B, H, W, _lib.ptr(w_k), _lib.ptr(bias), _lib.ptr(rp) if res else None,
                                              1, _lib.ptr(out), _lib.stream_ptr()))
    us_old = timed(lambda: L.cadre_conv2d_nhwc(_lib.ptr(xp), B, H, W, C, _lib.ptr(w_k), C, 3, 3, 1, 1, _lib.ptr(bias),
                                              _lib.ptr(r_old) if res else None, 0, 1, _lib.ptr(old), 1, _lib.stream_ptr()))
    print(f"B=640 res={res}: halo128 {us:.1f} us = {fl / us / 1e6:.0f} TFLOP/s   implicit GEMM {us_old:.1f} us = {fl / us_old / 1e6:.0f} TFLOP/s", flush=True)
