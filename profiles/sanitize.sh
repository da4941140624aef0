#!/bin/bash
# compute-sanitizer over the persistent kernels: bf16 train steps (fused attention chains forward + BPTT,
# decoder-LSTM chains) at the bench batch shape with 3 frames, then 2 inference steps.  Logs -> gpurun_out/sanitize_*.log
# (summaries are copied to profiles/).  Usage: bash profiles/sanitize.sh [memcheck racecheck synccheck ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in "${@:-memcheck synccheck racecheck}"; do
  for t in $tool; do
    echo "== $t"
    timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool "$t" --print-limit 20 python profiles/run_step.py 3 2 bf16 \
        > "gpurun_out/sanitize_$t.log" 2>&1
    echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok |Error|error" "gpurun_out/sanitize_$t.log" | head -8
  done
done
