#!/bin/bash
# compute-sanitizer passes over the small-case GPU parity tests (memcheck: out-of-bounds / misaligned accesses and
# API errors; racecheck: shared-memory hazards in the tiled / gather / pseudo-event kernels; initcheck: reads of
# uninitialised global memory, e.g. a workspace region a kernel assumed zeroed).  Run on a B200:
#   bash tools/sanitize.sh > gpurun_out/sanitizer.txt 2>&1
SEL='golden or randomised or modes_identical or abi_error or fused_augment or resize or mixed_image or from_timestamps'
for tool in ${TOOLS:-memcheck racecheck}; do      # TOOLS=initcheck is very slow here (>15 min): pick a small -k selection
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$SEL" 2>&1 | grep -v "^=========$" | tail -40
  echo "=== $tool exit: ${PIPESTATUS[0]}"
done
