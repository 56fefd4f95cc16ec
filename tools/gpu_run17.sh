#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 400 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | grep -vE "^\s*$|^\.+|^s\.+" | grep -E "assert|Error|passed|failed|FAILED|^E " | cut -c1-600 | tail -8
  echo "-- run $i done"
done
