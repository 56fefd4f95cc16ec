#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one plain step, (2) full capture of the dominant kernel, (3) attention
set -x
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01.csv python tools/one_step.py > gpurun_out/ncu_step.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_r01.csv > gpurun_out/launches_r01_summary.txt 2>&1
head -30 gpurun_out/launches_r01_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop_tc -c 3 -o gpurun_out/prof_conv_fprop_r01 \
    python tools/bench_conv.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 6 -o gpurun_out/prof_attn_r01 \
    python tools/bench_attn.py > gpurun_out/ncu_attn.log 2>&1
timeout 200 python tools/bench_attn.py > gpurun_out/bench_attn.log 2>&1; cat gpurun_out/bench_attn.log
ls -la gpurun_out/*.ncu-rep
