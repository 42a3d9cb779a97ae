"""Pinned host -> device bandwidth: one vs. two copy streams, 70 MB chunks (the bench's rgb chunk)."""
import torch, time
dev = "cuda:0"
n = 10
src = [torch.empty(70_778_880, dtype=torch.uint8, pin_memory=True) for _ in range(n)]
for t in src: t.fill_(1)
dst = [torch.empty(70_778_880, dtype=torch.uint8, device=dev) for _ in range(n)]
for ns in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(n):
            with torch.cuda.stream(streams[i % ns]):
                dst[i].copy_(src[i], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{ns} stream(s): {n * 70.78 / dt / 1e3:.1f} GB/s", flush=True)
# one big buffer
big = torch.empty(707_788_800, dtype=torch.uint8, pin_memory=True); big.fill_(1)
bigd = torch.empty(707_788_800, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); bigd.copy_(big, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"single 708 MB copy: {707.8 / dt / 1e3:.1f} GB/s")
