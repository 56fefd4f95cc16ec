#!/bin/bash
# operand-builder kernels of the attention node: parity tests, penalty step with / without, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "^E |assert|Error|passed|failed|FAILED" | cut -c1-400 | tail -14
GG_TIMING_ONLY=1 GG_ATTN_AUGMENT=0 timeout 300 python tools/profile_phases.py 2>&1 | grep "graph replay" | sed 's/^/[concat] /'
GG_TIMING_ONLY=1 timeout 300 python tools/profile_phases.py 2>&1 | grep "graph replay" | sed 's/^/[augment] /'
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-420; tail -2 gpurun_out/bench.err | cut -c1-300
