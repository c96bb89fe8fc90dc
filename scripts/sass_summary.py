"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA, TMA, TMEM loads, FP64 tensor
MMA, bulk copies, cluster ops) in the built library.  usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "macrodna_b200", "libmacrodna_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "DMMA", "HMMA", "IMMA", "SYNCS",
        "UCGABAR", "REDUX", "ATOMG", "RED\\.", "UBLKPF"]
cur = None
counts = collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + p, line):
            counts[cur][p.replace("\\.", "")] += 1
            if p.startswith("UTC") or p in ("UTMALDG", "UBLKCP", "UBLKPF") or (p == "ATOMG" and "CAS.128" in line):
                mm = re.search(r"\b(" + p + r"[.\w]*)", line)
                if mm:
                    counts[cur]["  " + mm.group(1)] += 1
print("# cuobjdump -sass macrodna_b200/libmacrodna_b200.so (sm_100a), mnemonic counts per kernel")
print("# built from git", subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip())
for fn, c in counts.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name).split("(")[0]
    keys = [k for k in c if not k.startswith("  ")]
    sub = [k for k in c if k.startswith("  ")]
    print("%-60s %s" % (name[:60], "  ".join("%s=%d" % (k, c[k]) for k in keys)))
    if sub:
        print(" " * 62 + "  ".join("%s=%d" % (k.strip(), c[k]) for k in sorted(sub)))
