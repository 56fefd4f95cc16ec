#!/bin/bash
# round 2, session 3, call 3: attention kernels with deeper key / query stages, attention node fixed
mkdir -p gpurun_out
timeout 300 python tools/bench_attn.py > gpurun_out/attn_ab2.jsonl 2> gpurun_out/attn_ab2.err
cat gpurun_out/attn_ab2.jsonl | cut -c1-330; tail -3 gpurun_out/attn_ab2.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/tests.log
tail -12 gpurun_out/tests.log
timeout 300 python tools/diag_fastpaths.py fp32 > gpurun_out/diag_fast_fp32.txt 2>&1
head -12 gpurun_out/diag_fast_fp32.txt
GG_ATTN_NODE=0 timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile_nonode.txt 2>&1
grep "graph replay" gpurun_out/phase_profile_nonode.txt
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
