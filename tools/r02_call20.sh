#!/bin/bash
# gather-shape sweep: phase times of the C2 step for each variant library (tools/build_variant.sh) and the shipped one
mkdir -p gpurun_out
{
python tools/phase_times.py --bins 5
python tools/phase_times.py --bins 1
python tools/phase_times.py --bins 1 --store soa
for so in cmda_b200/variants/lib_*.so; do
  CMDA_B200_LIB=$PWD/$so timeout 120 python tools/phase_times.py --bins 5
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/r02_gather_sweep.txt 2>&1
cat gpurun_out/r02_gather_sweep.txt
