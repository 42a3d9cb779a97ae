"""profiles/r2_*.md from the files a tools/gpu_round2_profile.sh run left in gpurun_out/ (read here, no GPU):
  r2_bench_launches.md  - ncu launch list of ONE resident bench step, grouped by kernel, shares vs bench.py's CUDA events
  r2_ppo_ncu_full.md    - ncu --set full of ONE PPO update step: per-kernel duration, tensor pipe, L2 / DRAM throughput,
                          DRAM bytes, registers, achieved occupancy
usage: python tools/make_r2_profiles.py <tag>"""
import collections, csv, io, json, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r"
bench = json.loads(open(f"gpurun_out/r2_bench_{tag}.json").read().strip().splitlines()[-1])

def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name).replace("cadre::", "").replace("void ", "")
    name = re.sub(r"\((int|bool|unsigned int)\)", "", name)
    return re.sub(r"\((?!.*>).*$", "", name) if ">" in name else re.sub(r"\(.*$", "", name)

# ---------------------------------------------------------------- launch list
rows = [r for r in csv.DictReader(l for l in open(f"gpurun_out/r2_bench_launches_{tag}.csv") if l.startswith('"'))]
agg = collections.OrderedDict()
for r in rows:
    k = short(r["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r["Metric Value"]) / 1e3
tot = sum(a[1] for a in agg.values())
enc_names = ("tc_persist", "tc_flat3x3", "tc_stem", "tc_gemm_kernel<0", "preprocess", "route_max", "pam_", "cam_", "intertask",
             "pack_f32", "f32_to_enc", "window_scatter")
enc = sum(a[1] for k, a in agg.items() if k.startswith(enc_names))
tc = sum(a[1] for k, a in agg.items() if k.startswith(("tc_", "lstm_seq")))
ph = bench["phases"]
L = [f"# r2: ncu launch list of ONE resident learner step of bench.py (1 GPU, cfg3: 6400 window frames + GAE + 8 PPO update steps), grouped by kernel",
     "# command: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv python bench.py --profile-step --warmup 2 ...",
     "# (tools/gpu_round2_profile.sh; the step sits between cudaProfilerStart/Stop). Per-launch times are cold-cache and serialised - the two",
     "# encoder streams and the PPO side stream do not overlap under ncu - so compare SHARES with bench.py's CUDA-event phases of the same box:",
     f"# bench.py ({tag}): {bench['ms_per_step']:.2f} ms/step = encoder {ph['encoder_ms']} ms + GAE {ph['gae_ms']} ms + PPO {ph['ppo_update_ms']} ms -> encoder share "
     f"{ph['encoder_ms'] / (ph['encoder_ms'] + ph['gae_ms'] + ph['ppo_update_ms']):.3f}",
     f"# launches in the step: {len(rows)}; sum of kernel durations: {tot / 1e3:.2f} ms", "kernel | launches | total us | share"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    L.append(f"{k} | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}%")
L.append(f"# tcgen05 kernels (tc_*, lstm_seq_*): {100 * tc / tot:.1f}% of the step; encoder-side kernels: {100 * enc / tot:.1f}% "
         f"(bench.py CUDA events: {100 * ph['encoder_ms'] / (ph['encoder_ms'] + ph['gae_ms'] + ph['ppo_update_ms']):.1f}%)")
open("profiles/r2_bench_launches.md", "w").write("\n".join(L) + "\n")

# ---------------------------------------------------------------- PPO ncu full
raw = subprocess.run(["ncu", "-i", f"gpurun_out/r2_ppo_full_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True,
                     text=True, check=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rd[0], rd[1], rd[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default
def f(r, name):
    try:
        return float(g(r, name).replace(",", ""))
    except ValueError:
        return float("nan")
def unit(name):
    return units[col[name]] if name in col else ""
M = {"dur": "gpu__time_duration.sum", "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "tensor2": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
     "lts": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "regs": "launch__registers_per_thread",
     "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "l1": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"}
def to_us(r):
    v, u = f(r, M["dur"]), unit(M["dur"])
    return v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
def to_mb(r, name):
    v, u = f(r, name), unit(name)
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    return v * scale
tensor_key = M["tensor"] if M["tensor"] in col else M["tensor2"]
P = ["# r2: ncu --set full of ONE PPO update step (tools/ncu_ppo.py: 4 workers x 100 rows x 2 heads, routed; cfg 3), every launch in",
     "# execution order (main stream and side stream interleaved as ncu serialised them). --clock-control none.",
     "# tensor% = " + tensor_key + "; L2% = lts__throughput; DRAM% = gpu__dram_throughput (pct of peak sustained elapsed).",
     f"# launches: {len(data)}; sum of durations: {sum(to_us(r) for r in data):.1f} us (serialised, cold caches; bench.py's CUDA-event update step: "
     f"{bench['rooflines'][0]['ms'] * 1e3:.0f} us fwd+bwd + {bench['rooflines'][1]['ms'] * 1e3:.0f} us clip+Adam)",
     "# | kernel | grid | block | us | tensor% | SM% | L1% | L2% | DRAM% | DRAM MB rd+wr | regs | warps active% |"]
for r in data:
    P.append(f"| {short(g(r, 'Kernel Name'))} | {g(r, 'Grid Size')} | {g(r, 'Block Size')} | {to_us(r):.1f} | {f(r, tensor_key):.1f} | {f(r, M['sm']):.1f} | "
             f"{f(r, M['l1']):.1f} | {f(r, M['lts']):.1f} | {f(r, M['dram']):.1f} | {to_mb(r, M['rd']) + to_mb(r, M['wr']):.1f} | {g(r, M['regs'])} | {f(r, M['occ']):.1f} |")
open("profiles/r2_ppo_ncu_full.md", "w").write("\n".join(P) + "\n")
print("\n".join(L[:12])); print("\n".join(P))
