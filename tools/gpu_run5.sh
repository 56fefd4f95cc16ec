#!/bin/bash
# ncu evidence + full bench line (both reference arms)
mkdir -p gpurun_out
bash tools/run_ncu.sh r02 > gpurun_out/run_ncu.log 2>&1
tail -25 gpurun_out/run_ncu.log
timeout 900 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-3000
