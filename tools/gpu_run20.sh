#!/bin/bash
# the driver's reference arm on the GPU box (CPU implementation of the reference on the host cores)
mkdir -p gpurun_out
nproc
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 8 --warmup 4 ) > gpurun_out/bench_refarm.log 2> gpurun_out/bench_refarm.err
tail -1 gpurun_out/bench_refarm.log | cut -c1-900
tail -4 gpurun_out/bench_refarm.err
