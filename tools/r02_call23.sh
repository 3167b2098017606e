#!/bin/bash
mkdir -p gpurun_out
SEL='golden or randomised or modes_identical or abi_error or fused_augment or resize or mixed_image or from_timestamps or side_stream or packed_store'
{
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$SEL" 2>&1 | grep -v "^=========$" | tail -25
  echo "=== $tool exit: ${PIPESTATUS[0]}"
done
} > gpurun_out/r02_sanitizer.txt 2>&1
tail -30 gpurun_out/r02_sanitizer.txt
