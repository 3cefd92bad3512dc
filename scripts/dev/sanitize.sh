#!/bin/bash
# compute-sanitizer over a small-shape subset of the hot path: memcheck, racecheck (shared-memory hazards of the
# byte-counter histograms, ring buffers, ticket hand-overs), synccheck, initcheck.  Only kernels of namespace `upr` are checked.
# Run on the GPU box:   gpurun -- 'bash scripts/dev/sanitize.sh'      Logs: gpurun_out/sanitize_<tool>.log (summaries -> profiles/)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/../..}"
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for tool in memcheck racecheck synccheck initcheck; do
    extra=""
    [ "$tool" = "memcheck" ] && extra="--leak-check no"
    [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
    timeout "${SANITIZE_TIMEOUT:-900}" compute-sanitizer --tool "$tool" $extra --kernel-name kns=upr --print-limit 40 \
        --log-file "gpurun_out/sanitize_${tool}.log" python scripts/dev/sanitize_subset.py > "gpurun_out/sanitize_${tool}.out" 2>&1
    echo "$tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "gpurun_out/sanitize_${tool}.log" | tail -1) : $(tail -1 gpurun_out/sanitize_${tool}.out)"
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
    # the hand-rolled P2P flag protocol of the fused statistics + exchange kernel, two ranks
    timeout 600 compute-sanitizer --tool synccheck --kernel-name kns=upr --target-processes all --log-file gpurun_out/sanitize_peer_synccheck.%p.log \
        python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/peer_allreduce_check.py \
        > gpurun_out/sanitize_peer.out 2>&1
    echo "peer synccheck rc=$? : $(grep -h 'ERROR SUMMARY' gpurun_out/sanitize_peer_synccheck.*.log | sort | uniq -c | tr '\n' ' ')"
fi
