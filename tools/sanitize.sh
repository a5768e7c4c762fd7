#!/bin/bash
# compute-sanitizer over small invocations of the hot path.  Run on a GPU box:
#   gpurun -- 'bash tools/sanitize.sh'      -> gpurun_out/sanitize_*.log ; summaries are copied to profiles/ by hand
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, stage, extra env
  local tool=$1 stage=$2
  ( time timeout -s KILL 400 $CS --tool $tool --print-limit 20 python tools/sanitize_target.py $stage ) > gpurun_out/sanitize_${tool}_${stage}.log 2>&1
  echo "== $tool $stage: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|ok' gpurun_out/sanitize_${tool}_${stage}.log | tr '\n' ' ')"
}
run memcheck mc
run memcheck sampler
run synccheck sampler
run racecheck mc
SAN_STEPS=2 run racecheck sampler
