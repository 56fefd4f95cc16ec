#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_bmm.py > gpurun_out/bmm_node.jsonl 2>&1
cat gpurun_out/bmm_node.jsonl
