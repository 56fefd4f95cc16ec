#!/bin/bash
# round 2, session 3, call 4: coalesced thin-layer loader, attention warp-count defaults, ncu source captures of the thin kernels
mkdir -p gpurun_out
timeout 300 python tools/bench_layers.py > gpurun_out/conv_layers.jsonl 2>&1
grep -E "D256|D128|G128|G256|total" gpurun_out/conv_layers.jsonl | cut -c1-250
timeout 300 python tools/bench_attn.py > gpurun_out/attn_ab3.jsonl 2> gpurun_out/attn_ab3.err
cut -c1-200 gpurun_out/attn_ab3.jsonl
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/tests.log
tail -5 gpurun_out/tests.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_thin_tc_kernel -c 1 -o gpurun_out/prof_thin_fprop_r02 \
    python tools/bench_layers.py D256_block2 > gpurun_out/ncu_thin1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_thin_wgrad -c 1 -o gpurun_out/prof_thin_wgrad_r02 \
    python tools/bench_layers.py D256_block2 > gpurun_out/ncu_thin2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn2_ -c 3 -o gpurun_out/prof_attn2b_r02 \
    python tools/bench_attn.py D_res32_l2 0 > gpurun_out/ncu_attn2b.log 2>&1
ls -la gpurun_out/*.ncu-rep
