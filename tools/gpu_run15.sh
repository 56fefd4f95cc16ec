#!/bin/bash
# round 2, session 3: unit-coefficient axpby gradients passed through; smoke(); full GPU suite; bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/tests.log
tail -4 gpurun_out/tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -1 gpurun_out/bench.log | cut -c1-420
timeout 300 python tools/profile_phases.py > gpurun_out/phase_profile.txt 2>&1
grep "graph replay" gpurun_out/phase_profile.txt
grep -E "axpby_kernel<__nv_bfloat16>" gpurun_out/phase_profile.txt | cut -c1-110
