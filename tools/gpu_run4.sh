#!/bin/bash
# full GPU test-suite + per-phase profile + bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v "^    " | tail -40 > gpurun_out/tests.log
tail -6 gpurun_out/tests.log
timeout 500 python tools/profile_phases.py 256 16 > gpurun_out/phases.log 2>&1
grep "graph replay" gpurun_out/phases.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
