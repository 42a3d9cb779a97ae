"""torchrun probe of the in-switch all-reduce (cadre_b200.collective.SwitchAllReduce) against NCCL: result check
(fp64 sum of the per-rank inputs, bit-identical replicas) and device time for the learner's message sizes."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: E402
from cadre_b200.collective import SwitchAllReduce  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    __graft_entry__.build()
dist.barrier()
N = 19470272
out = {"world": world}
for mc in ([True, False] if os.environ.get("PROBE_P2P", "1") == "1" else [True]):
    ar = SwitchAllReduce(N, dev, multicast=mc)
    tag = "multicast" if ar.multicast else "p2p"
    if mc and not ar.multicast:
        out["multicast"] = "unavailable"
        continue
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(N, device=dev, generator=g)
    # reference: gather every rank's input, sum in fp64
    ref = torch.zeros(N, device=dev, dtype=torch.float64)
    for r in range(world):
        gr = torch.Generator(device=dev).manual_seed(100 + r)
        ref += torch.randn(N, device=dev, generator=gr).double()
    res = {}
    for off, cnt in ((0, N), (1024, 9000000), (N - 4096, 4096)):
        ar.buffer.copy_(x)
        torch.cuda.synchronize(); dist.barrier()
        ar.sum_(off, cnt)
        ar.check()
        got = ar.buffer[off:off + cnt].double()
        err = ((got - ref[off:off + cnt]).abs().max() / ref.abs().max()).item()
        untouched = bool(torch.equal(ar.buffer[:off], x[:off]) and torch.equal(ar.buffer[off + cnt:], x[off + cnt:]))
        # replicas bit-identical: compare with rank 0's bits
        mine = ar.buffer[off:off + cnt].clone()
        dist.broadcast(mine, 0)
        same = bool(torch.equal(mine, ar.buffer[off:off + cnt]))
        res[f"range_{off}_{cnt}"] = {"max_err_rel": err, "untouched_outside": untouched, "replicas_identical": same}
    out[tag + "_check"] = res
    sizes = {"full_77.9MB": (0, N), "range_36MB": (0, 9005760), "small_1.4MB": (N - 345600, 345600)}
    for blocks in [int(b) for b in os.environ.get("PROBE_BLOCKS", "32,64,128,192").split(",")]:
        ar.set_blocks(blocks)
        for name, (off, cnt) in sizes.items():
            for _ in range(5):
                ar.sum_(off, cnt)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ar.sum_(off, cnt)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            out[f"{tag}_b{blocks}_{name}"] = {"ms": round(ms.item(), 4), "algbw_GBps": round(cnt * 4 / ms.item() / 1e6, 1)}
    ar.check()
    del ar
# NCCL on the same sizes
y = torch.randn(N, device=dev)
for name, (off, cnt) in {"full_77.9MB": (0, N), "range_36MB": (0, 9005760), "small_1.4MB": (N - 345600, 345600)}.items():
    for _ in range(5):
        dist.all_reduce(y[off:off + cnt])
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(y[off:off + cnt])
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    out[f"nccl_{name}"] = {"ms": round(ms.item(), 4), "algbw_GBps": round(cnt * 4 / ms.item() / 1e6, 1)}
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
