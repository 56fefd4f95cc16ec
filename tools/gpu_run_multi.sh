#!/bin/bash
# N-GPU batch: NCCL gradient-equivalence test, full single-GPU suite, bench lines of every BASELINE config at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/multi_smi.txt
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "^    " | tail -40 > gpurun_out/tests_multi.log
  tail -6 gpurun_out/tests_multi.log
fi
run() {  # config steps warmup
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29507 \
      bench.py --gpus $N --config $1 --steps $2 --warmup $3 --no-cpu-baseline --no-gpu-reference \
      > gpurun_out/bench_$1_n$N.log 2> gpurun_out/bench_$1_n$N.err
  tail -1 gpurun_out/bench_$1_n$N.log | cut -c1-900
}
run cfg2 8 4
run cfg3 4 4
run cfg4 4 4
run cfg5 4 4
