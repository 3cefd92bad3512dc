#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out/r3_red2.log
: > $O
timeout 300 python -m pytest tests/test_clahe_gpu.py -x -q -m gpu >> $O 2>&1
for v in 8 12 0 8 12 0; do
  UPR_CLAHE_VARIANT=$v timeout 120 python scripts/stage_bench.py >> $O 2>&1
done
timeout 200 python scripts/fused_bench.py >> $O 2>&1
tail -30 $O
