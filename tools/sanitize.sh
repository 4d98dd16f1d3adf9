#!/bin/bash
# compute-sanitizer over small-N runs of the hot kernels (GPU box): memcheck + racecheck + synccheck; summaries to gpurun_out/
cd ${GRAFT_REPO_ROOT:-.}
for t in ff gather9 gather32; do
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $t > gpurun_out/sanitize_${t}_${tool}.log 2>&1
    echo "== $t $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' gpurun_out/sanitize_${t}_${tool}.log | tr '\n' ' ')"
  done
done
