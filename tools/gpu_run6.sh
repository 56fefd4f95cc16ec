#!/bin/bash
# round 2, session 3, call 1: attention kernel generations A/B (+ deviation), GPU parity tests, bench line (no reference arms)
mkdir -p gpurun_out
timeout 300 python tools/bench_attn.py > gpurun_out/attn_ab.jsonl 2> gpurun_out/attn_ab.err
cat gpurun_out/attn_ab.jsonl; tail -5 gpurun_out/attn_ab.err
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests.log
tail -6 gpurun_out/tests.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-1200
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
