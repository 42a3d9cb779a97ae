"""bench.py — CADRE learner hot path on B200: encoder forward + GAE + PPO update (+ gradient all-reduce + Adam).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cadre|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

One STEP = one learner iteration of BASELINE config 3 on every rank (weak scaling: per-GPU work is fixed):
  4 logical workers x T=200 env steps x 8 stacked frames -> encoder forward over 6400 camera frames (uint8 rgb +
  route map + measurements), features scattered into the 8 RolloutStorages, GAE + advantage normalisation,
  then ppo_epoch=4 x 2 minibatches of (routed LSTM/actor-critic forward + backward, gradient all-reduce over
  NCCL for N > 1, per-module clip + Adam).
value   = frames/s with inputs resident in HBM (device-timed, max over ranks, whole job).
e2e     = the same through the public API with HOST inputs: pinned uint8 frames copied H2D every step inside the
          timed region (chunked, overlapped with the encoder on a copy stream) and the losses read back D2H.
--impl reference times the reference's own CPU implementation (oracle port: /root/reference is not on the GPU
box) on a bounded sample of the same workload with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKERS, T, SEQ, MB_NUM, PPO_EPOCH = 4, 200, 8, 2, 4
FRAMES_PER_STEP = WORKERS * T * SEQ           # per rank
ENC_CHUNK = int(os.environ.get("CADRE_BENCH_CHUNK", "640"))   # frames per encoder call (divides 6400)
METRIC = "encoder+PPO-update frames/sec"
WORKLOAD = ("cfg3 learner iteration per GPU: 4 workers x T=200 x 8 frames encoder fwd (6400 u8 frames 144x256) + "
            "GAE + 4 epochs x 2 minibatches x 400 rows PPO update (fwd+bwd, allreduce, clip+Adam)")


# ---------------------------------------------------------------------------------------------- helpers
def encoder_flops_per_frame():
    """name -> FLOPs/frame of each encoder launch as EXECUTED here (SURVEY.md §8d counts 3.0875 GFLOP/frame for
    the unfused reference graph; the folded fc1 does less work)."""
    f = {}
    f["stem"] = f["stem+pool"] = 2 * 72 * 128 * 64 * 196
    H, W, C = 36, 64, 64
    for li, planes in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            s = 2 if (li > 1 and bi == 0) else 1
            Ho = (H + 2 - 3) // s + 1
            Wo = (W + 2 - 3) // s + 1
            f[f"layer{li}.{bi}.conv1"] = 2 * Ho * Wo * planes * C * 9
            f[f"layer{li}.{bi}.conv2"] = 2 * Ho * Wo * planes * planes * 9
            if s == 2:
                f[f"layer{li}.0.downsample"] = 2 * Ho * Wo * planes * C
                f[f"layer{li}.0.conv2+shortcut"] = f[f"layer{li}.0.conv2"] + f[f"layer{li}.0.downsample"]
            H, W, C = Ho, Wo, planes
    f["conv5a|conv5c"] = 2 * 40 * 256 * 512 * 9
    f["conv51"] = f["conv52+sum"] = 2 * 40 * 128 * 128 * 9
    f["fc1(folded)"] = 2 * 5120 * 3072
    f["fc2"] = 2 * 6 * 512 * 256
    f["pam.value_conv"] = 2 * 40 * 128 * 128
    return f


def sample_clocks_start():
    path = tempfile.mktemp(prefix="cadre_clocks_", suffix=".csv")
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                              "-i", os.environ.get("LOCAL_RANK", "0")], stdout=open(path, "w"),
                             stderr=subprocess.DEVNULL)
    except Exception:
        return None, path
    return p, path


def sample_clocks_stop(p, path):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    if p is None:
        return out
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    sm, mx, reasons = [], [], set()
    try:
        for line in open(path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(path)
    except Exception:
        pass
    if sm:
        busy = sorted(sm)[len(sm) // 4:]          # drop idle samples at the edges
        out["sm_mhz"] = float(np.median(busy))
        out["sm_max_mhz"] = float(max(mx))
        out["reasons"] = sorted(reasons)
    return out


def ncu_traffic():
    """DRAM bytes (read + write) of the tcgen05 launches of one encoder forward at batch 640, from the committed
    `ncu --set full` capture (profiles/r1_final_encoder_ncu_full.md); None if the summary is missing."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["tcgen05_family_dram_bytes_per_forward_b640"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1417.3), d.get("hbm_gbs", 6453.1), "measured"
    return 1400.0, 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------- reference arm
def cpu_reference_rate(n_frames, mb, threads):
    """frames/s of a full learner step on the CPU, extrapolated from separately timed bounded samples:
    t_step = 6400 * t_enc_per_frame + 4 * t_gae_pair + 3200 * t_update_policy_per_row + 8 * t_chief_step."""
    from oracle import restate as R
    torch.set_num_threads(threads)
    rs = np.random.RandomState(0)
    sd = R.danet_fixture_state(0)
    psd = R.ppo_fixture_state(0)
    params = {m: {n: t.clone().requires_grad_(True) for n, t in d.items()} for m, d in psd.items()}
    adam = {m: {n: {"exp_avg": torch.zeros_like(t), "exp_avg_sq": torch.zeros_like(t)} for n, t in d.items()}
            for m, d in psd.items()}
    ticks = [R.synthetic_tick(rs) for _ in range(max(1, n_frames // 8))]
    sts = [R.synthetic_storage(rs, T=T, actions=a) for a in (R.STEER_ACTIONS, R.THROTTLE_ACTIONS)]
    t0 = time.perf_counter()
    with torch.no_grad():
        for tk in ticks:
            R.agent_latent_feature(tk["rgb"], tk["route_fig"].copy(), tk["measurements"], sd)
    t_enc = (time.perf_counter() - t0) / (len(ticks) * 8)
    t0 = time.perf_counter()
    advs = []
    for st in sts:
        st["returns"], vp = R.compute_returns(st["rewards"], st["value_preds"], st["masks"], torch.tensor([[0.1]]))
        advs.append(R.normalized_advantages(st["returns"], vp))
    t_gae = time.perf_counter() - t0
    idx = [list(range(mb)), list(range(mb))]
    samples = [R.gather_minibatch(st, adv, ix) for st, adv, ix in zip(sts, advs, idx)]
    t0 = time.perf_counter()
    R.update_policy(samples[0], samples[1], params)
    t_upd = (time.perf_counter() - t0) / mb
    grads = {m: {n: p.grad for n, p in d.items()} for m, d in params.items()}
    t0 = time.perf_counter()
    R.chief_step(params, grads, adam, step=1)
    t_chief = time.perf_counter() - t0
    n_upd = PPO_EPOCH * MB_NUM
    rows_per_step = n_upd * WORKERS * (T // MB_NUM)
    t_step = FRAMES_PER_STEP * t_enc + WORKERS * t_gae + rows_per_step * t_upd + n_upd * t_chief
    return FRAMES_PER_STEP / t_step, dict(t_enc_per_frame_ms=t_enc * 1e3, t_gae_pair_ms=t_gae * 1e3,
                                          t_update_per_row_ms=t_upd * 1e3, t_chief_step_ms=t_chief * 1e3,
                                          t_step_s=t_step)


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU BEFORE the pinned staging buffers
    are allocated (first-touch puts their pages on that NUMA node). With 8 ranks copying 944 MB per step each, the
    host-to-device leg of `e2e` is bound by host memory / inter-socket bandwidth, not by the GPUs."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        orig = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in orig]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus), orig
    except Exception:
        pass
    return None, None


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    detail = None
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_rate(32, 32, threads)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * FRAMES_PER_STEP / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": "per step: encoder on 32 frames + GAE on 2 storages + update_policy/chief on a "
                                   "32-row minibatch (dense 4-command compute), extrapolated to the 6400-frame / "
                                   "3200-row step", **detail},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- cadre arm
def run_cadre(args):
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the "
                         "CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    numa, full_affinity = bind_to_gpu_numa_node(local)
    dist = torch.distributed
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from cadre_b200.encoder import Encoder
    from cadre_b200.learner import Learner, RolloutPool
    from cadre_b200 import fixtures as R   # seeded synthetic weights (no checkpoint on the box)

    enc = Encoder(R.danet_fixture_state(0), dev, max_batch=ENC_CHUNK)
    mb = T // MB_NUM
    learner = Learner(WORKERS, mb, R.ppo_fixture_state(0), dev, seeds=[rank * WORKERS + w for w in range(WORKERS)],
                      process_group=None)
    pool = RolloutPool(WORKERS, dict(num_steps=T, mini_batch_num=MB_NUM, feature_dims=530, seq_length=SEQ,
                                     use_gae=True, gamma=0.99, tau=0.95), dev)
    # synthetic rollout: frames + per-step scalars, seeded per rank
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    n = FRAMES_PER_STEP
    rgb = torch.randint(0, 256, (n, 144, 256, 3), device=dev, dtype=torch.uint8, generator=g)
    route = (torch.rand(n, 256, 144, device=dev, generator=g) < 0.1).to(torch.uint8) * 255
    meas = torch.rand(n, 3, device=dev, dtype=torch.float64, generator=g)
    b = pool.batched
    b["rewards"].copy_(torch.rand(b["rewards"].shape, device=dev, generator=g))
    b["masks"].copy_((torch.rand(b["masks"].shape, device=dev, generator=g) >= 0.02).float())
    b["command"].copy_(torch.randint(0, 4, b["command"].shape, device=dev, generator=g, dtype=torch.int32))
    b["action_log_probs"].copy_(-0.5 - 2.5 * torch.rand(b["action_log_probs"].shape, device=dev, generator=g))
    b["value_preds"].copy_(torch.randn(b["value_preds"].shape, device=dev, generator=g))
    for w in range(WORKERS):
        b["action"][2 * w].copy_(torch.randint(0, 33, (T + 1, 1), device=dev, generator=g))
        b["action"][2 * w + 1].copy_(torch.randint(0, 3, (T + 1, 1), device=dev, generator=g))
    next_values = torch.zeros(WORKERS, 2, device=dev)
    feats = torch.empty(n, 530, device=dev)

    def scatter_features():
        # features of (worker, step, frame) -> obs[t] of both heads' storages (train.py:69-72 inserts the same
        # obs_feature into the steer and the throttle rollout)
        f = feats.view(WORKERS, T, SEQ, 530)
        for w in range(WORKERS):
            b["obs"][2 * w, :T].copy_(f[w])
            b["obs"][2 * w + 1, :T].copy_(f[w])

    # The encoder chunks are independent: with two encoder instances on two streams the tail of one chunk's
    # persistent kernels (SMs idle while the last tiles finish) is filled by the other chunk's CTAs.
    NS = max(1, int(os.environ.get("CADRE_ENC_STREAMS", "2")))
    encs = [enc] + [Encoder(R.danet_fixture_state(0), dev, max_batch=ENC_CHUNK) for _ in range(NS - 1)]
    enc_streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    fork_ev = torch.cuda.Event()
    join_ev = [torch.cuda.Event() for _ in range(NS)]

    def encode_chunks(chunk_inputs):
        """chunk_inputs: iterable of (k, prepare) where prepare(stream) returns (rgb, route, meas, out) for chunk k"""
        main = torch.cuda.current_stream()
        if NS == 1:
            for i, prep in chunk_inputs:
                a = prep(main)
                enc.forward_u8(*a)
            return
        fork_ev.record(main)
        for st in enc_streams:
            st.wait_event(fork_ev)
        for i, prep in chunk_inputs:
            st = enc_streams[i % NS]
            with torch.cuda.stream(st):
                a = prep(st)
                encs[i % NS].forward_u8(*a)
        for k, st in enumerate(enc_streams):
            join_ev[k].record(st)
            main.wait_event(join_ev[k])

    def step_resident():
        encode_chunks([(i, (lambda st, s=s: (rgb[s:s + ENC_CHUNK], route[s:s + ENC_CHUNK], meas[s:s + ENC_CHUNK],
                                             feats[s:s + ENC_CHUNK])))
                       for i, s in enumerate(range(0, n, ENC_CHUNK))])
        scatter_features()
        pool.compute_returns(next_values)
        return learner.learn(pool, PPO_EPOCH)

    # host-resident variant: pinned inputs, H2D on a copy stream double-buffered against the encoder
    rgb_h = torch.empty(rgb.shape, dtype=torch.uint8, pin_memory=True).copy_(rgb)
    route_h = torch.empty(route.shape, dtype=torch.uint8, pin_memory=True).copy_(route)
    meas_h = torch.empty(meas.shape, dtype=torch.float64, pin_memory=True).copy_(meas)
    copy_stream = torch.cuda.Stream(device=dev)
    NB = 2 * NS   # staging buffers: two per encoder stream
    stage = [(torch.empty_like(rgb[:ENC_CHUNK]), torch.empty_like(route[:ENC_CHUNK]), torch.empty_like(meas[:ENC_CHUNK]))
             for _ in range(NB)]
    staged_ev = [torch.cuda.Event() for _ in range(NB)]
    free_ev = [torch.cuda.Event() for _ in range(NB)]
    losses_h = torch.empty(WORKERS, 2, 3, pin_memory=True)
    h2d_bytes = rgb_h.numel() + route_h.numel() + meas_h.numel() * 8
    d2h_bytes = losses_h.numel() * 4

    # e2e chunk schedule: the encoder cannot start before its first chunk has landed, so the first chunks are
    # small (128 frames = 19 MB = 0.35 ms of PCIe time instead of 1.7 ms for 640 frames)
    if ENC_CHUNK % 640 == 0 and n % ENC_CHUNK == 0 and n > ENC_CHUNK:
        ramp = [128, 512] + [640 * 2 ** i for i in range(8) if 640 * 2 ** (i + 1) <= ENC_CHUNK]   # 128, 512, 640, 1280, ...
        e2e_sizes = ramp + [ENC_CHUNK] * ((n - sum(ramp)) // ENC_CHUNK)
    else:
        e2e_sizes = [ENC_CHUNK] * (n // ENC_CHUNK)
    assert sum(e2e_sizes) == n
    e2e_starts = [sum(e2e_sizes[:i]) for i in range(len(e2e_sizes))]

    def step_e2e():
        def prep_for(i, s, m):
            def prep(st):
                k = i % NB
                with torch.cuda.stream(copy_stream):
                    if i >= NB:
                        copy_stream.wait_event(free_ev[k])
                    stage[k][0][:m].copy_(rgb_h[s:s + m], non_blocking=True)
                    stage[k][1][:m].copy_(route_h[s:s + m], non_blocking=True)
                    stage[k][2][:m].copy_(meas_h[s:s + m], non_blocking=True)
                    staged_ev[k].record(copy_stream)
                st.wait_event(staged_ev[k])
                return stage[k][0][:m], stage[k][1][:m], stage[k][2][:m], feats[s:s + m]
            return prep

        chunks = list(enumerate(zip(e2e_starts, e2e_sizes)))
        main = torch.cuda.current_stream()
        fork_ev.record(main)
        copy_stream.wait_event(fork_ev)   # the previous step's consumers of the staging buffers are done
        if NS > 1:
            for st in enc_streams:
                st.wait_event(fork_ev)
        for i, (s, m) in chunks:
            st = enc_streams[i % NS] if NS > 1 else main
            with torch.cuda.stream(st):
                a = prep_for(i, s, m)(st)
                encs[i % NS].forward_u8(*a)
                free_ev[i % NB].record(st)
        if NS > 1:
            for k, st in enumerate(enc_streams):
                join_ev[k].record(st)
                main.wait_event(join_ev[k])
        scatter_features()
        pool.compute_returns(next_values)
        learner.learn(pool, PPO_EPOCH)
        losses_h.copy_(learner.losses, non_blocking=True)

    def timed(fn, steps, warmup, sample=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        proc = path = None
        if sample:
            proc, path = sample_clocks_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sample_clocks_stop(proc, path) if sample else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps, clocks

    if world > 1:
        learner.world = world   # all-reduce inside update_step
    ms_step, clocks = timed(step_resident, args.steps, args.warmup, sample=True)
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup - 1))
    total_frames = FRAMES_PER_STEP * world
    value = total_frames / (ms_step * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)

    # ---- phase split of one resident step (every rank runs it: update_step contains the all-reduce)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for s in range(0, n, ENC_CHUNK):
        enc.forward_u8(rgb[s:s + ENC_CHUNK], route[s:s + ENC_CHUNK], meas[s:s + ENC_CHUNK], feats[s:s + ENC_CHUNK])
    ev[1].record()
    scatter_features()
    pool.compute_returns(next_values)
    ev[2].record()
    n_upd = learner.learn(pool, PPO_EPOCH)
    ev[3].record()
    torch.cuda.synchronize()
    phase = {"encoder_ms": round(ev[0].elapsed_time(ev[1]), 3), "scatter_gae_ms": round(ev[1].elapsed_time(ev[2]), 3),
             "ppo_update_ms": round(ev[2].elapsed_time(ev[3]), 3), "update_steps": n_upd,
             "encoder_frames_per_s": round(n / (ev[0].elapsed_time(ev[1]) * 1e-3)),
             "ppo_samples_per_s": round(WORKERS * T * PPO_EPOCH / (ev[2].elapsed_time(ev[3]) * 1e-3))}

    # ---- the one exchange step of the path: all-reduce(sum) of the flat fp32 gradient, timed alone
    if world > 1:
        dist.barrier()
        for _ in range(3):
            dist.all_reduce(learner.grads, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        ar0, ar1 = torch.cuda.Event(True), torch.cuda.Event(True)
        ar0.record()
        for _ in range(10):
            dist.all_reduce(learner.grads, op=dist.ReduceOp.SUM)
        ar1.record()
        torch.cuda.synchronize()
        ar = torch.tensor([ar0.elapsed_time(ar1) / 10], device=dev)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        nbytes = learner.grads.numel() * 4
        phase["allreduce"] = {"bytes": nbytes, "ms": round(float(ar.item()), 4),
                              "algbw_GBps": round(nbytes / (float(ar.item()) * 1e-3) / 1e9, 1),
                              "busbw_GBps": round(2 * (world - 1) / world * nbytes / (float(ar.item()) * 1e-3) / 1e9, 1),
                              "per_step": n_upd}
        learner.grads.zero_()

    # ---- per-kernel view (rank 0): encoder launches timed with CUDA events inside the library
    roofline, kernels = None, None
    if rank == 0:
        tf_peak, hbm_peak, peak_src = measured_peaks()
        enc.forward_u8(rgb[:ENC_CHUNK], route[:ENC_CHUNK], meas[:ENC_CHUNK], feats[:ENC_CHUNK])
        prof = None
        for _ in range(3):
            prof = enc.profile(ENC_CHUNK)
        fl = encoder_flops_per_frame()
        kernels = []
        for name, ms in prof:
            ent = {"name": name, "ms": round(ms, 4)}
            if name in fl:
                ent["tflops"] = round(fl[name] * ENC_CHUNK / (ms * 1e-3) / 1e12, 1)
            kernels.append(ent)
        conv = [k for k in kernels if "tflops" in k]
        tot_ms = sum(k["ms"] for k in conv)
        tot_fl = sum(fl[k["name"]] for k in conv) * ENC_CHUNK
        ach = tot_fl / (tot_ms * 1e-3) / 1e12
        roofline = {"kernel": "tcgen05 tile kernels (implicit-GEMM conv / linear launches of one encoder forward, "
                              f"{len(conv)} launches, batch {ENC_CHUNK})",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": round(ach / tf_peak, 4), "traffic": ncu_traffic(), "peak_source": peak_src + " (sustained bf16)",
                    "share_of_encoder_ms": round(tot_ms / sum(k["ms"] for k in kernels), 3)}
        launches = (n // ENC_CHUNK) * enc.launches_per_forward + 1 + n_upd * (learner.engine.launches + 3)
        if full_affinity:
            os.sched_setaffinity(0, full_affinity)   # the CPU baseline uses every host core
        cpu_val, cpu_detail = cpu_reference_rate(64, 64, os.cpu_count() or 1)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate (encoder), tf32 / fp32 (PPO)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "workers_per_gpu": WORKERS, "num_steps": T, "seq_length": SEQ,
                       "mini_batch": mb, "ppo_epoch": PPO_EPOCH, "encoder_chunk": ENC_CHUNK, "encoder_streams": NS,
                       "l2": "inputs (944 MB of uint8 frames per step) and activations exceed the 126 MB L2",
                       "host_cpus_bound_per_rank": numa},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches * args.steps),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "encoder on 64 frames + GAE on 2 storages + update_policy/chief on a 64-row "
                                       "minibatch, extrapolated to the 6400-frame / 3200-row step", **cpu_detail},
            "phases": phase, "kernels": kernels,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cadre", choices=["cadre", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_cadre(args)


if __name__ == "__main__":
    main()
