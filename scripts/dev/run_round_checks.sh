#!/bin/bash
# The command sequence behind profiles/r4_*: GPU tests, smoke, the default bench line, the reference arm, the ncu launch list and the
# two ncu --set full captures.   gpurun --timeout 2400 -- 'bash scripts/dev/run_round_checks.sh'
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r4_tests.log 2>&1; echo "tests rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r4_bench_n1.json 2> gpurun_out/r4_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-named > gpurun_out/r4_ncu_launches.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_hist_lab_vec3|k_map_vec5" -c 3 -s 0 -o gpurun_out/r4_clahe_full -f python scripts/stage_bench.py > gpurun_out/r4_ncu_full.log 2>&1; echo "clahe full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_saliency_stream2|k_sal_normalize|k_att_gain|k_ms_stream" -c 4 -s 4 -o gpurun_out/r4_ca_full -f python scripts/dev/ca_once.py > gpurun_out/r4_ncu_ca.log 2>&1; echo "content-aware full rc=$?"
tail -3 gpurun_out/r4_tests.log; cut -c1-400 gpurun_out/r4_bench_n1.json
