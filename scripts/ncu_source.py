#!/usr/bin/env python3
"""Per-instruction digest of an ncu report's source page (SASS view) for one kernel launch.

    python scripts/ncu_source.py REPORT.ncu-rep KERNEL [--skip N] [--top 25]

Prints: total stall samples by reason, shared-memory wavefronts (ideal vs actual) by opcode, and the hottest instructions.
"""
import argparse
import collections
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("kernel")
ap.add_argument("--skip", type=int, default=0)
ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", a.kernel, "--launch-skip", str(a.skip),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
def num(r, k):
    try:
        return float(r.get(k, 0) or 0)
    except ValueError:
        return 0.0
stall_keys = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = collections.Counter()
for r in rows:
    for k in stall_keys:
        tot[k] += num(r, k)
allsmp = sum(tot.values())
print("stall samples (all):", ", ".join(f"{k[6:]} {v / allsmp * 100:.1f}%" for k, v in tot.most_common(10)))
inst = sum(num(r, "Instructions Executed") for r in rows)
print(f"warp instructions executed: {inst:.0f}")
wf = collections.Counter(); wfi = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if num(r, "L1 Wavefronts Shared") > 0:
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        wf[op] += num(r, "L1 Wavefronts Shared"); wfi[op] += num(r, "L1 Wavefronts Shared Ideal"); cnt[op] += num(r, "Instructions Executed")
print("shared wavefronts by opcode (actual / ideal / warp-instr):")
for op, v in wf.most_common():
    print(f"  {op:12s} {v:14.0f} {wfi[op]:14.0f} {cnt[op]:14.0f}  x{v / max(cnt[op], 1):.2f} per instr")
print(f"  total {sum(wf.values()):.0f} wavefronts, ideal {sum(wfi.values()):.0f}")
print("hottest instructions by stall samples:")
for r in sorted(rows, key=lambda r: -num(r, "# Samples"))[:a.top]:
    st = sorted(((num(r, k), k[6:]) for k in stall_keys), reverse=True)[:2]
    print(f"  {num(r, '# Samples'):7.0f}  {r['Source'].strip()[:70]:70s} " + " ".join(f"{k}:{v:.0f}" for v, k in st) +
          (f"  wf {num(r, 'L1 Wavefronts Shared'):.0f}/{num(r, 'L1 Wavefronts Shared Ideal'):.0f}" if num(r, "L1 Wavefronts Shared") else ""))
