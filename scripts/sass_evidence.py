#!/usr/bin/env python3
"""SASS evidence for the hardware features the kernels claim (no GPU needed): per kernel, the count of the mnemonics that prove
them in the shipped library.   python scripts/sass_evidence.py > profiles/r4_sass.md"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "retinex-image-enhancement_b200/libupretinex_b200.so"
WATCH = ["FFMA2", "FADD2", "FMUL2", "UTMALDG", "SYNCS", "RED", "ATOMS", "ATOMG", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "CCTL", "PRMT",
         "MUFU", "SHFL", "BAR", "VIMNMX", "HADD2", "F2FP", "DADD"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print(f"# SASS evidence (`cuobjdump -sass {LIB}`, sm_100a)\n")
print("Counts of the mnemonics that matter, per kernel (static instructions).  `FFMA2`/`FADD2`/`FMUL2` are Blackwell's packed fp32\n"
      "instructions; `UTMALDG` + `SYNCS` are TMA bulk-tensor loads completing on an mbarrier (extension Gaussian); `ATOMS` in the\n"
      "histogram kernel is `red.shared.add.u32`, the one-instruction byte-counter update; the 128-bit LDG/STG columns are the vectorised HBM streams.\n")
rows = []
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    ops = collections.Counter()
    total = 0
    for line in blk.split("\n"):
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            total += 1
            op = m.group(1)
            for w in WATCH:
                if "." in w:      # e.g. LDG.E.128: any LDG with a .128 qualifier (cache hints sit in between)
                    base, width = w.split(".")[0], w.split(".")[-1]
                    if op.split(".")[0] == base and "." + width in op:
                        ops[w] += 1
                elif op == w or op.startswith(w + "."):
                    ops[w] += 1
    rows.append((dem, total, ops))
hdr = ["kernel", "instr"] + WATCH
print("| " + " | ".join(hdr) + " |")
print("|" + "---|" * len(hdr))
for dem, total, ops in sorted(rows):
    if not dem.startswith(("void upr::", "upr::")):
        continue
    print("| `" + dem.replace("void ", "").replace("upr::", "") + "` | " + str(total) + " | " + " | ".join(str(ops.get(w, 0) or "") for w in WATCH) + " |")
