#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/phase_times.py --bins 5
for so in cmda_b200/variants/lib_*.so; do
  CMDA_B200_LIB=$PWD/$so timeout 120 python tools/phase_times.py --bins 5
done
} > gpurun_out/r02_call24.txt 2>&1
cat gpurun_out/r02_call24.txt
