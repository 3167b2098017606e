#!/bin/bash
mkdir -p gpurun_out
{
for so in shipped cmda_b200/variants/lib_oldvf.so cmda_b200/variants/lib_curvf.so; do
  if [ $so = shipped ]; then timeout 120 python tools/phase_times.py --bins 1 --store soa; else CMDA_B200_LIB=$PWD/$so timeout 120 python tools/phase_times.py --bins 1 --store soa; fi
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"band_|fallback" -c 12 python tools/phase_times.py --bins 1 --store soa --steps 2 2>&1 | grep -E "band_|fallback|gpu__time" | head -40
} > gpurun_out/r02_call28.txt 2>&1
cat gpurun_out/r02_call28.txt
