#!/bin/bash
# final state: GPU suite, smoke, bench line with both reference arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -E "^E |assert|Error|passed|failed|FAILED" | cut -c1-300 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
timeout 900 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-500
