#!/bin/bash
# round 2, session 3, call 2: fp32 fast-path diagnostic, GPU tests (all), penalty step with/without the attention node,
# ncu source-level capture of the second-generation attention kernels
mkdir -p gpurun_out
timeout 300 python tools/diag_fastpaths.py fp32 > gpurun_out/diag_fast_fp32.txt 2>&1
cat gpurun_out/diag_fast_fp32.txt | tail -48
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/tests.log
tail -12 gpurun_out/tests.log
GG_ATTN_NODE=0 timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile_nonode.txt 2>&1
grep "graph replay" gpurun_out/phase_profile_nonode.txt
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_ -c 3 -o gpurun_out/prof_attn2_r02 \
    python tools/bench_attn.py D_res32_l2 0 > gpurun_out/ncu_attn2.log 2>&1
ncu -i gpurun_out/prof_attn2_r02.ncu-rep --page raw --csv > gpurun_out/prof_attn2_r02.raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
