#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -k "attention or ka5 or ka8 or readme256 or ka2 or trainer" 2>&1 | grep -E "^E |assert|Error|passed|failed|FAILED" | cut -c1-300 | tail -6
GG_TIMING_ONLY=1 timeout 300 python tools/profile_phases.py 2>&1 | grep "graph replay"
