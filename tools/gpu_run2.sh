#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_parity.py > gpurun_out/diag.log 2>&1
cat gpurun_out/diag.log | cut -c1-400 | tail -40
timeout 900 python -m pytest tests/test_gpu_readme256.py -m gpu -q --timeout 600 -k "ka8 or aux_recon" --tb=short 2>&1 | grep -v "^    " > gpurun_out/tests_fail.log
tail -5 gpurun_out/tests_fail.log
