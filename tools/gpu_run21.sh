#!/bin/bash
# per-GPU batches of the strong-scaling rows at N = 4 and N = 8 (batch 4 and 2), run on one GPU: same shapes, no NCCL
mkdir -p gpurun_out
for b in 2 4; do
  timeout 300 python bench.py --batch $b --steps 4 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_b$b.log 2> gpurun_out/bench_b$b.err
  tail -1 gpurun_out/bench_b$b.log | cut -c1-260; tail -2 gpurun_out/bench_b$b.err | cut -c1-300
done
for c in cfg3 cfg4 cfg5; do
  timeout 300 python bench.py --config $c --steps 4 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err
  tail -1 gpurun_out/bench_$c.log | cut -c1-260; tail -2 gpurun_out/bench_$c.err | cut -c1-300
done
