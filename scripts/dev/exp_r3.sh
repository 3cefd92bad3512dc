#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r3_labmod.log
: > $O
for m in 0 16 8 4 2; do
  if [ $m = 0 ]; then timeout 120 python scripts/stage_bench.py >> $O 2>&1; else UPR_LAB_MOD=$m timeout 120 python scripts/stage_bench.py >> $O 2>&1; fi
done
tail -40 $O
