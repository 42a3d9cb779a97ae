"""All-reduce(sum) of the flat fp32 gradient (77.9 MB) and of one pipeline range (36 MB) over NCCL, for the
environment this process was launched with (NCCL_ALGO / NCCL_PROTO / ...). torchrun --nproc-per-node N."""
import os, sys, json
import torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
out = {}
for name, n in (("full_77.9MB", 19470272), ("range_36MB", 9039680 + 350000), ("small_1.4MB", 350000)):
    x = torch.randn(n, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[name] = {"ms": round(ms.item(), 4), "algbw_GBps": round(n * 4 / ms.item() / 1e6, 1),
                 "busbw_GBps": round(2 * (world - 1) / world * n * 4 / ms.item() / 1e6, 1)}
if dist.get_rank() == 0:
    print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}, "world": world, **out}), flush=True)
dist.barrier(); dist.destroy_process_group()
