#!/bin/bash
# round-2 first GPU batch: parity tests, microbenchmark, per-phase profile, per-layer table, bench line (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 100 python tools/bench_mma_chain.py > gpurun_out/mma_chain.txt 2>&1
timeout 500 python tools/profile_phases.py 256 16 > gpurun_out/phases.log 2>&1
grep "graph replay" gpurun_out/phases.log
timeout 300 python tools/bench_layers.py > gpurun_out/conv_layers_main.jsonl 2>&1
tail -1 gpurun_out/conv_layers_main.jsonl
timeout 100 python tools/bench_attn.py > gpurun_out/attn_main.jsonl 2>&1
cat gpurun_out/attn_main.jsonl
timeout 700 python bench.py --steps 8 --warmup 4 > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-3000
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.log | cut -c1-1500
