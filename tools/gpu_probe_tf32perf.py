"""Micro-benchmarks of the PPO-shaped GEMMs (why is a 17-k-block TF32 tile so slow?)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
L = _lib.lib(); dev = "cuda:0"

def timed(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

def gemm(kind, M, N, K, batch, lda=None, ldb=None, a_mn=0, b_mn=0, block_n=0, name=""):
    dt = torch.float32 if kind else _lib.enc_dtype()
    lda = lda or ((K + 3) // 4 * 4 if not a_mn else (M + 3) // 4 * 4)
    ldb = ldb or ((K + 3) // 4 * 4 if not b_mn else (N + 3) // 4 * 4)
    rows_a = K if a_mn else M
    rows_b = K if b_mn else N
    A = torch.randn(batch, rows_a, lda, device=dev).to(dt)
    B = torch.randn(batch, rows_b, ldb, device=dev).to(dt)
    out = torch.empty(batch, M, N + (4 - N % 4) % 4, device=dev)
    g = _lib.GemmArgs()
    g.kind, g.a_mn, g.b_mn, g.batch, g.M, g.N, g.K, g.block_n = kind, a_mn, b_mn, batch, M, N, K, block_n
    g.A, g.B, g.lda, g.a_bs, g.ldb, g.b_bs = A.data_ptr(), B.data_ptr(), lda, rows_a * lda, ldb, rows_b * ldb
    g.out, g.ldc, g.out_bs, g.out_f32, g.alpha = out.data_ptr(), out.shape[2], M * out.shape[2], 1, 1.0
    sp = _lib.stream_ptr()
    rc = L.cadre_gemm(ctypes.byref(g), sp)
    if rc: print(name, "ERR", L.cadre_last_error().decode()); return
    us = timed(lambda: L.cadre_gemm(ctypes.byref(g), sp))
    print(f"{name:55s} {us:8.1f} us  {2*M*N*K*batch/us/1e6:8.1f} TFLOP/s", flush=True)

gemm(1, 128, 2120, 530, 8, name="tf32 KK 128x2120x530 b8 (lstm step shape)")
gemm(1, 128, 2120, 530, 8, lda=9 * 532, name="tf32 KK same, lda=9*532 (strided rows)")
gemm(1, 128, 2120, 544, 8, lda=544, ldb=544, name="tf32 KK K=544 ld=544 (128B-aligned rows)")
gemm(1, 128, 2048, 512, 8, name="tf32 KK 128x2048x512 b8 (aligned)")
gemm(0, 128, 2120, 536, 8, lda=536, ldb=536, name="fp16 KK 128x2120x536 b8")
gemm(1, 128, 2120, 530, 1, name="tf32 KK 128x2120x530 b1")
gemm(1, 1152, 2120, 530, 8, name="tf32 KK 1152x2120x530 b8 (xpart shape)")
gemm(1, 128, 530, 2120, 8, b_mn=1, ldb=532, name="tf32 K/MN 128x530x2120 b8 (dgrad shape)")
gemm(1, 2120, 530, 900, 8, a_mn=1, b_mn=1, lda=2120, ldb=532, name="tf32 MN/MN 2120x530x900 b8 (wgrad shape)")
gemm(1, 4096, 4096, 4096, 1, name="tf32 KK 4096^3")

# ---- host enqueue cost vs device time
import time
A = torch.randn(8, 128, 532, device=dev); B = torch.randn(8, 2120, 532, device=dev); out = torch.empty(8, 128, 2120, device=dev)
g = _lib.GemmArgs()
g.kind, g.batch, g.M, g.N, g.K = 1, 8, 128, 2120, 530
g.A, g.B, g.lda, g.a_bs, g.ldb, g.b_bs = A.data_ptr(), B.data_ptr(), 532, 128 * 532, 532, 2120 * 532
g.out, g.ldc, g.out_bs, g.out_f32, g.alpha = out.data_ptr(), 2120, 128 * 2120, 1, 1.0
sp = _lib.stream_ptr()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): L.cadre_gemm(ctypes.byref(g), sp)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e6*(t1-t0)/200:.1f} us/call; drain after enqueue {1e6*(t2-t1):.1f} us total")
# device time of one launch measured in a CUDA graph of 50 launches
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    spp = ctypes.c_void_p(s.cuda_stream)
    L.cadre_gemm(ctypes.byref(g), spp)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s):
        for _ in range(50): L.cadre_gemm(ctypes.byref(g), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
gr.replay(); torch.cuda.synchronize()
e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
print(f"device time per launch inside a CUDA graph: {e0.elapsed_time(e1)/50*1e3:.1f} us")
