#!/bin/bash
mkdir -p gpurun_out
for w in gp plain g; do
  timeout 300 python tools/profile_aten_shapes.py $w > gpurun_out/aten_$w.txt 2>&1
  tail -48 gpurun_out/aten_$w.txt | cut -c1-240
done
