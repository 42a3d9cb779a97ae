"""Two launches of cadre_conv3x3_flat128 and of the implicit-GEMM kernel at B = 640 for ncu (-k regex:halo128|tc_persist)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
L = _lib.lib(); dt = _lib.enc_dtype(); dev = "cuda:0"
B, H, W, C = 640, 18, 32, 128
xp = torch.randn(B, H + 2, W + 2, C, device=dev).to(dt)
w_k = (torch.randn(C, 9 * C, device=dev) / 34.0).to(dt)
bias = torch.randn(C, device=dev)
out = torch.empty(B, H + 2, W + 2, C, device=dev, dtype=dt)
old = torch.empty(B, H, W, C, device=dev, dtype=dt)
for _ in range(2):
    L.cadre_conv3x3_flat128(_lib.ptr(xp), B, H, W, _lib.ptr(w_k), _lib.ptr(bias), None, 1, _lib.ptr(out), _lib.stream_ptr())
    L.cadre_conv2d_nhwc(_lib.ptr(xp), B, H, W, C, _lib.ptr(w_k), C, 3, 3, 1, 1, _lib.ptr(bias), None, 0, 1, _lib.ptr(old), 1, _lib.stream_ptr())
torch.cuda.synchronize()
