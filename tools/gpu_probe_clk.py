import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cadre_b200 import _lib
dev = "cuda:0"
clk = torch.zeros(17 * 8 * 8, dtype=torch.int64, device=dev)
os.environ["CADRE_DBG_CLK"] = str(clk.data_ptr())
L = _lib.lib()
A = torch.randn(8, 128, 532, device=dev); B = torch.randn(8, 2120, 532, device=dev); out = torch.empty(8, 128, 2120, device=dev)
g = _lib.GemmArgs()
g.kind, g.batch, g.M, g.N, g.K = 1, 8, 128, 2120, 530
g.A, g.B, g.lda, g.a_bs, g.ldb, g.b_bs = A.data_ptr(), B.data_ptr(), 532, 128 * 532, 532, 2120 * 532
g.out, g.ldc, g.out_bs, g.out_f32, g.alpha = out.data_ptr(), 2120, 128 * 2120, 1, 1.0
for _ in range(3):
    L.cadre_gemm(ctypes.byref(g), _lib.stream_ptr())
torch.cuda.synchronize()
c = clk.view(-1, 8).cpu()
d = c[:, :8] - c[:, :1]
names = ["start", "after alloc+sync", "first full (MMA)", "MMA all issued", "epilogue sees tmem_full", "epilogue done", "after dealloc", "phase 1 done"]
import statistics
for i, n in enumerate(names):
    col = d[:, i].tolist()
    print(f"{n:28s} median {statistics.median(col):9.0f} cyc  min {min(col):9.0f} max {max(col):9.0f}")
# back-to-back timing of the same launch (no stamps)
g2 = _lib.GemmArgs.from_buffer_copy(g)
os.environ.pop("CADRE_DBG_CLK", None)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
for name, gg in (("stamped", g),):
    for _ in range(5):
        L.cadre_gemm(ctypes.byref(gg), _lib.stream_ptr())
    e0.record()
    for _ in range(50):
        L.cadre_gemm(ctypes.byref(gg), _lib.stream_ptr())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"{name}: {us:.1f} us per launch, weights {8*2120*532*4/1e6:.1f} MB -> {8*2120*532*4/us/1e6:.2f} TB/s, {2*8*128*2120*530/us/1e6:.1f} TFLOP/s")
