#!/bin/bash
# final bench line of the session with both reference arms + phase profile
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err
tail -1 gpurun_out/bench_full.log | cut -c1-600
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
