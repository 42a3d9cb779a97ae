"""Experiment: does a UMMA smem descriptor whose start address is shifted by r rows (r*128 B, not 1024-aligned)
read rows r.. of a TMA-written SWIZZLE_128B tile correctly, and does it need base_offset = r?"""
import ctypes, os, sys, subprocess
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
L = _lib.lib()
dev = "cuda:0"
shift = int(os.environ.get("CADRE_DBG_A_SHIFT", "0"))
torch.manual_seed(0)
M, N, K = 128, 128, 64
A = torch.randn(M, K, device=dev).to(_lib.enc_dtype())
B = torch.randn(N, K, device=dev).to(_lib.enc_dtype())
out = torch.zeros(M, N, device=dev)
g = _lib.GemmArgs()
g.kind, g.batch, g.M, g.N, g.K = 0, 1, M, N, K
g.A, g.B, g.lda, g.ldb = A.data_ptr(), B.data_ptr(), K, K
g.out, g.ldc, g.out_f32, g.alpha = out.data_ptr(), N, 1, 1.0
rc = L.cadre_gemm(ctypes.byref(g), _lib.stream_ptr())
torch.cuda.synchronize()
ref = A.float() @ B.float().t()
n = M - shift
err = (out[:n] - ref[shift:shift + n]).abs().max().item()
err0 = (out[:n] - ref[:n]).abs().max().item()
print(f"shift={shift} base_offset={os.environ.get('CADRE_DBG_BASE_OFFSET','0')} rc={rc} "
      f"err_vs_shifted={err:.3e} err_vs_unshifted={err0:.3e} ref_max={ref.abs().max().item():.2f}")
