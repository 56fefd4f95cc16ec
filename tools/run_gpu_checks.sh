#!/bin/bash
# Standard GPU check batch (one gpurun call, ~3 min on a normal box):
#   parity tests, per-layer conv table, attention bench, thin-kernel timeline, step profile, bench line.
# Usage: gpurun --timeout 1500 -- 'bash tools/run_gpu_checks.sh [quick]'
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
if [ "$1" != "quick" ]; then
  timeout 300 python tools/bench_layers.py > gpurun_out/conv_layers.jsonl 2>&1
  tail -1 gpurun_out/conv_layers.jsonl
  timeout 100 python tools/bench_attn.py > gpurun_out/attn.jsonl 2>&1
  cat gpurun_out/attn.jsonl
  timeout 100 python tools/trace_thin.py 16 32 > gpurun_out/trace_thin.txt 2>&1
  tail -4 gpurun_out/trace_thin.txt
  timeout 300 python tools/profile_step.py 256 16 > gpurun_out/profile_step.log 2>&1
  grep -E "^\[|Self CUDA time total" gpurun_out/profile_step.log
fi
timeout 300 python bench.py --steps 8 --warmup 8 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-1600
