#!/bin/bash
# standard GPU check batch: parity tests, conv microbench, step timing (eager + graphs)
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/tests.log
tail -3 gpurun_out/tests.log
timeout 600 python tools/profile_step.py 256 16 > gpurun_out/profile_step.log 2>&1
grep -E "^\[" gpurun_out/profile_step.log
timeout 900 python bench.py --steps 8 --warmup 8 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log
