"""SASS evidence per kernel of libcadre_sm100.so (runs here, no GPU): counts of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, LDGMC/STGMC-class = multimem, HMMA = mma.sync).
usage: python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cadre_b200", "libcadre_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTC?MMA other", r"\bUTC[A-GI-Z]MMA"),
       ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("UBLKCP", r"\bUBLKCP"), ("LDGMC (multimem.ld_reduce)", r"\bLDGMC"), 
       ("HMMA (mma.sync)", r"\bHMMA"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("SYNCS (mbarrier)", r"\bSYNCS")]
cur, counts, lines = None, collections.OrderedDict(), collections.Counter()
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in ln:
        continue
    lines[cur] += 1
    for name, pat in PAT:
        if re.search(pat, ln):
            counts[cur][name] += 1
def demangle(n):
    out = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    out = re.sub(r"^void ", "", out)
    out = re.sub(r"\((int|bool|unsigned int)\)", "", out)
    out = re.sub(r"\((?!.*>).*$", "", out) if ">" in out else re.sub(r"\(.*$", "", out)
    out = out.replace("<unnamed>::", "")
    return out.replace("cadre::", "")
print("# r2: SASS mnemonics per kernel of cadre_b200/libcadre_sm100.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG, multimem.ld_reduce -> LDGMC (multimem.st is a plain STG.E.128.STRONG.SYS to the multicast address)")
cols = [n for n, _ in PAT]
print("kernel | SASS lines | " + " | ".join(cols))
tot = collections.Counter()
for k, c in counts.items():
    if not any(c.values()) and lines[k] < 50:
        continue
    print(demangle(k) + f" | {lines[k]} | " + " | ".join(str(c.get(n, 0)) for n in cols))
    tot.update(c)
print("TOTAL | " + str(sum(lines.values())) + " | " + " | ".join(str(tot.get(n, 0)) for n in cols))
