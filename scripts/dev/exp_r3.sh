#!/bin/bash
cd "$(dirname "$0")/../.."
for F in 2 4 8; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -k regex:k_ --csv --log-file gpurun_out/r3_l2probe_F$F.csv python scripts/dev/l2_probe.py $F > gpurun_out/r3_l2probe_F$F.log 2>&1
done
tail -3 gpurun_out/r3_l2probe_F4.log
