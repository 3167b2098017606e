#!/bin/bash
# side-stream plans + 128-thread gather variants + the GPU suite
mkdir -p gpurun_out
{
timeout 120 python tools/phase_times.py --bins 5
timeout 120 python tools/phase_times.py --bins 1
timeout 120 python tools/phase_times.py --bins 1 --store soa
for so in cmda_b200/variants/lib_*.so; do
  CMDA_B200_LIB=$PWD/$so timeout 120 python tools/phase_times.py --bins 5
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/r02_call21.txt 2>&1
cat gpurun_out/r02_call21.txt
