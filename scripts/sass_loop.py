#!/usr/bin/env python3
"""Offline SASS accounting (no GPU needed): instruction mix of the innermost hot loop of a kernel.

    python scripts/sass_loop.py k_map_vec5 [--lib path.so] [--px 4] [--dump]

Finds every backward branch in the kernel's SASS, takes the loop with the most instructions (or --loop N) and prints
its opcode histogram and instructions per pixel (--px = pixels processed per loop iteration per thread).
"""
import argparse
import collections
import re
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("kernel")
ap.add_argument("--lib", default="retinex-image-enhancement_b200/libupretinex_b200.so")
ap.add_argument("--px", type=float, default=4)
ap.add_argument("--loop", type=int, default=None)
ap.add_argument("--dump", action="store_true")
a = ap.parse_args()

sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
blk = [b for b in blocks if a.kernel in b.split("\n", 1)[0]]
assert blk, f"kernel {a.kernel} not found"
ins = []
for line in blk[0].split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        loops.append((int(m.group(1), 16), addr))
print(f"{blk[0].splitlines()[0].strip()}: {len(ins)} SASS instructions, loops: " +
      ", ".join(f"[{lo:#x},{hi:#x}]={(hi - lo) // 16 + 1}" for lo, hi in loops))
if not loops:
    raise SystemExit
lo, hi = max(loops, key=lambda l: l[1] - l[0]) if a.loop is None else loops[a.loop]
# innermost = the largest loop that contains no other loop? take the requested / largest one
body = [(ad, t) for ad, t in ins if lo <= ad <= hi]
ops = collections.Counter()
for _, t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    ops[t.split()[0].split(".")[0]] += 1
n = len(body)
print(f"loop [{lo:#x},{hi:#x}]: {n} instructions / iteration = {n / a.px:.1f} per pixel")
print("  " + ", ".join(f"{k} {v} ({v / a.px:.1f})" for k, v in ops.most_common()))
if a.dump:
    for ad, t in body:
        print(f"  {ad:#06x}  {t}")
