#!/bin/bash
# round 2, session 3, call 5: per-warp |k|^2 staging, single-pass shared-QK L2 forward, warp-count defaults
mkdir -p gpurun_out
timeout 300 python tools/bench_attn.py > gpurun_out/attn_ab4.jsonl 2> gpurun_out/attn_ab4.err
cut -c1-215 gpurun_out/attn_ab4.jsonl; tail -3 gpurun_out/attn_ab4.err
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
