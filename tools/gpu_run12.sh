#!/bin/bash
# round 2, session 3: attention forward with two CTAs per SM; shared-bank prep channels per block sweep
mkdir -p gpurun_out
timeout 300 python tools/bench_attn.py > gpurun_out/attn_ab5.jsonl 2> gpurun_out/attn_ab5.err
grep -E "gen2_occ2|gen2_default" gpurun_out/attn_ab5.jsonl | cut -c1-215; tail -3 gpurun_out/attn_ab5.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fused_attention" 2>&1 | tail -4
for ob in 1 2 4; do
  echo "== GG_SB_OB=$ob"
  GG_SB_OB=$ob timeout 300 python tools/profile_phases.py > gpurun_out/phase_ob$ob.txt 2>&1
  grep "graph replay" gpurun_out/phase_ob$ob.txt | head -3
  grep "sbank_prep_kernel" gpurun_out/phase_ob$ob.txt | head -3 | cut -c1-100
done
echo "== GG_FLAGS=128"
GG_FLAGS=128 timeout 300 python tools/profile_phases.py > gpurun_out/phase_occ2.txt 2>&1
grep "graph replay" gpurun_out/phase_occ2.txt | head -3
