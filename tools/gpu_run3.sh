#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 --tb=short -k "shared_bank or ka1 or ka4 or cuda_graph_replay or ka6 or ka7 or ka9 or patch" 2>&1 | grep -v "^    " | tail -30 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 500 python tools/profile_phases.py 256 16 > gpurun_out/phases.log 2>&1
grep "graph replay" gpurun_out/phases.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
