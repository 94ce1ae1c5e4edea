#!/bin/bash
# plain-C host checks of the newest C-ABI entry points (no Python start-up): run on a GPU box, log under gpurun_out/
mkdir -p gpurun_out
{
  nvidia-smi --query-gpu=name --format=csv,noheader
  for d in "0,0" "0" "0,0,0"; do timeout 60 ./tests/c/test_multi_match "$d"; echo "rc=$? (multi_match $d)"; done
  timeout 60 ./tests/c/test_synth; echo "rc=$? (synth)"
} > gpurun_out/r02_c_checks.log 2>&1
tail -40 gpurun_out/r02_c_checks.log
