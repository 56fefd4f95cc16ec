#!/bin/bash
# ncu evidence for profiles/: (1) launch list (+DRAM bytes) of one plain training step, (2) full captures of the
# dominant kernels.  Numbers printed by runs under ncu are never bench values.
set -x
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/launches_r01.csv python tools/one_step.py > gpurun_out/ncu_step.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_r01.csv > gpurun_out/launches_r01_summary.txt 2>&1
head -40 gpurun_out/launches_r01_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop_tc -c 3 -o gpurun_out/prof_conv_fprop_r01 \
    python tools/bench_conv.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_thin_tc -c 2 -o gpurun_out/prof_conv_thin_r01 \
    python tools/bench_layers.py D256_block > gpurun_out/ncu_thin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 6 -o gpurun_out/prof_attn_r01 \
    python tools/bench_attn.py > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
