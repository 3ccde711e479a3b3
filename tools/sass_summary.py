"""Writes profiles/<round>_sass_summary.txt: SASS mnemonic counts per kernel of the built library (cuobjdump -sass), the evidence that
the contraction kernels are tcgen05 / TMEM / TMA code (B200_PROFILING.md names the mnemonics).  python tools/sass_summary.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "neuro__b200", "libneuro_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
keys = ["UTCHMMA", "2CTA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "LDGSTS", "FFMA", "LDG", "STG", "LDS", "STS", "SHFL"]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    c = collections.Counter()
    for i in ins:
        c[i.split(".")[0]] += 1
        if i.startswith("UTCHMMA") and ".2CTA" in i:
            c["2CTA"] += 1
    rows.append((name, len(ins), c))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.split("\n")
lines = ["# SASS mnemonic counts per kernel of neuro__b200/libneuro_b200.so (cuobjdump -sass, sm_100a); regenerate: python tools/sass_summary.py",
         "# UTCHMMA = tcgen05.mma kind::tf32 (2CTA = of which cta_group::2), LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG = TMA tensor load,",
         "# UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, LDGSTS = cp.async. No HMMA / IMMA (mma.sync) anywhere: see the last line.",
         "%-72s %6s " % ("kernel", "instrs") + " ".join("%7s" % k for k in keys)]
tot = collections.Counter()
for (n, t, c), d in sorted(zip(rows, names), key=lambda r: r[1]):
    d = re.sub(r"nb200::\(anonymous namespace\)::", "", d); d = re.sub(r"\(.*", "", d).replace("void ", "")
    lines.append("%-72s %6d " % (d[:72], t) + " ".join("%7d" % c[k] for k in keys))
    tot.update(c)
lines.append("# library totals: " + ", ".join("%s %d" % (k, tot[k]) for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA", "IMMA", "DMMA")))
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(l for l in lines if "tc_" in l or l.startswith("#") or l.startswith("kernel")))
