#!/usr/bin/env python
"""Instruction census of the product library's sm_100a SASS (cuobjdump -sass): the Blackwell-native mnemonics per kernel family.
    python tools/sass_census.py [tag]   ->  profiles/<tag>_sass_census.md
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce, UBLKCP = bulk copy,
HMMA / IMMA = legacy mma.sync (cross-check attention kernel, INT8 weight-streaming kernel), MUFU.EX2 = exp2."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "vit-fpga_b200", "lib", "libnetcuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "UTCIMMA", "UTCIMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "SYNCS",
        "USETMAXREG", "HMMA", "IMMA", "MUFU.EX2", "FMNMX3", "FFMA2", "FADD2", "LDGSTS"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("nc::", "")
        fam = re.sub(r"<.*", "", name)
        cur = per.setdefault(fam, collections.Counter())
        cur["__kernels"] += 1
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["__instrs"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                key = w + (".2CTA" if ".2CTA" in op and not w.endswith("2CTA") and w in ("UTCHMMA", "UTCIMMA") else "")
                cur[key] += 1
cols = [w for w in WANT if w not in ("UTCHMMA.2CTA", "UTCIMMA.2CTA")]
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCIMMA", "UTCIMMA.2CTA"] + [c for c in cols if c not in ("UTCHMMA", "UTCIMMA")]
out = [f"# SASS instruction census ({tag})", "",
       "`cuobjdump -sass vit-fpga_b200/lib/libnetcuda.so`, instructions per kernel family (all template instantiations summed).  " + __doc__.split("\n", 3)[3].strip().replace("\n", " "), "",
       "| kernel family | instantiations | SASS instructions | " + " | ".join(cols) + " |", "|---|---:|---:|" + "---:|" * len(cols)]
tot = collections.Counter()
for fam, c in per.items():
    out.append(f"| `{fam}` | {c['__kernels']} | {c['__instrs']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")
    tot.update(c)
out.append(f"| **total** | {tot['__kernels']} | {tot['__instrs']} | " + " | ".join(str(tot[k]) if tot[k] else "" for k in cols) + " |")
open(os.path.join(ROOT, "profiles", f"{tag}_sass_census.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
