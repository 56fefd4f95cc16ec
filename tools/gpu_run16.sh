#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cuda_graph_replay_matches_eager" 2>&1 | grep -E "assert|passed|failed|Error" | cut -c1-400 | tail -6
done
echo "== GG_AXPBY_PASS=0"
for i in 1 2; do
  GG_AXPBY_PASS=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cuda_graph_replay_matches_eager" 2>&1 | grep -E "assert|passed|failed|Error" | cut -c1-400 | tail -6
done
