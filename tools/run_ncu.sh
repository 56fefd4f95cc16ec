#!/bin/bash
# ncu evidence for profiles/ (round 2): (1) launch list (+DRAM bytes) of one plain training step, (2) full captures of
# the dominant kernels.  Numbers printed by runs under ncu are never bench values.
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/launches_$R.csv python tools/one_step.py > gpurun_out/ncu_step.log 2>&1
python tools/ncu_summarize.py gpurun_out/launches_$R.csv gpurun_out/${R}_ncu_step_launches.json > gpurun_out/launches_${R}_summary.txt 2>&1
head -30 gpurun_out/launches_${R}_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_fprop_tc -c 3 -o gpurun_out/prof_conv_fprop_$R \
    python tools/bench_conv.py > gpurun_out/ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn2_ -c 3 -o gpurun_out/prof_attn_$R \
    python tools/bench_attn.py D_res32_l2 0 > gpurun_out/ncu_attn.log 2>&1
for f in prof_conv_fprop_$R prof_attn_$R; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
