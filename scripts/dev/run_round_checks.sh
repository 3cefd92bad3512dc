#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -x -q -m gpu > gpurun_out/r3_tests.log 2>&1; echo "tests rc=$?" 
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 50 --warmup 5 > gpurun_out/r3_bench_c2.json 2> gpurun_out/r3_bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3_bench_ref.json 2> gpurun_out/r3_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r3_ncu_launches.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_hist_lab_vec3|k_map_vec5" -c 2 -s 0 -o gpurun_out/r3_clahe_full -f python scripts/stage_bench.py > gpurun_out/r3_ncu_full.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/r3_tests.log; cat gpurun_out/r3_bench_c2.json | cut -c1-400
