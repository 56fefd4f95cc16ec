#!/bin/bash
# round-2 first GPU batch: parity tests, microbenchmarks, per-phase profile, A/B of the r2-wip kernels, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 100 python tools/bench_mma_chain.py > gpurun_out/mma_chain.txt 2>&1
timeout 400 python tools/profile_phases.py 256 16 > gpurun_out/phases.log 2>&1
grep "graph replay" gpurun_out/phases.log
GG_NO_MERGE=1 GG_TIMING_ONLY=1 timeout 300 python tools/profile_phases.py 256 16 > gpurun_out/phases_nomerge.log 2>&1
grep "graph replay" gpurun_out/phases_nomerge.log
timeout 300 python tools/bench_layers.py > gpurun_out/conv_layers_main.jsonl 2>&1
tail -1 gpurun_out/conv_layers_main.jsonl
GG_LIB=$PWD/gigagan_pytorch_b200/libgigagan_sm100_wip.so timeout 300 python tools/bench_layers.py > gpurun_out/conv_layers_wip.jsonl 2>&1
tail -1 gpurun_out/conv_layers_wip.jsonl
timeout 100 python tools/bench_attn.py > gpurun_out/attn_main.jsonl 2>&1
GG_LIB=$PWD/gigagan_pytorch_b200/libgigagan_sm100_wip.so timeout 100 python tools/bench_attn.py > gpurun_out/attn_wip.jsonl 2>&1
cat gpurun_out/attn_main.jsonl gpurun_out/attn_wip.jsonl
GG_LIB=$PWD/gigagan_pytorch_b200/libgigagan_sm100_wip.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_readme256.py -m gpu -q -k "conv or attention or attn or ka8" 2>&1 | tail -8 > gpurun_out/tests_wip.log
tail -3 gpurun_out/tests_wip.log
GG_LIB=$PWD/gigagan_pytorch_b200/libgigagan_sm100_wip.so GG_TIMING_ONLY=1 timeout 300 python tools/profile_phases.py 256 16 > gpurun_out/phases_wip.log 2>&1
grep "graph replay" gpurun_out/phases_wip.log
timeout 600 python bench.py --steps 8 --warmup 4 > gpurun_out/bench.log 2>gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-3000
