#!/bin/bash
# run every GPU test function in its own process (a CUDA fault in one does not poison the others)
export CUDA_LAUNCH_BLOCKING=1
for t in $(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
  echo "=== $t"
  python -m pytest "$t" -x -q 2>&1 | grep -E "passed|failed|error|Error|^E  |^tests/.*:[0-9]+:|^subphaser_b200/.*:[0-9]+:" | head -${LINES_PER_TEST:-14}
done
