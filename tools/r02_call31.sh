#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_part3_b1 -k regex:"band_partition3|band_accumulate" -c 2 \
    python tools/profile_step.py --bins 1 --mode auto --steps 1 --store soa > gpurun_out/r02_part3_b1.log 2>&1
ls -la gpurun_out/
