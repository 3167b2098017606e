#!/bin/bash
# gather-shape sweep: phase times of the C2 step for each variant library (tools/build_variant.sh) and the shipped one
mkdir -p gpurun_out
{
timeout 120 python tools/phase_times.py --bins 5
timeout 120 python tools/phase_times.py --bins 1
timeout 120 python tools/phase_times.py --bins 1 --store soa
timeout 120 python tools/phase_times.py --bins 3
timeout 120 python tools/phase_times.py --bins 2
CMDA_B200_LIB=$PWD/cmda_b200/variants/lib_v5b1g4.so timeout 120 python tools/phase_times.py --bins 1
for so in cmda_b200/variants/lib_*.so; do
  CMDA_B200_LIB=$PWD/$so timeout 120 python tools/phase_times.py --bins 5
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/r02_gather_sweep.txt 2>&1
cat gpurun_out/r02_gather_sweep.txt
