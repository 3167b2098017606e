#!/bin/bash
# round 2, GPU call 1: split probe / banded phase times + ncu --set full of the pseudo-event kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools/round2_first_call.sh > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo \
    -k regex:"pair_|isr_" -c 12 python tools/profile_pseudo.py > gpurun_out/r02_pseudo_ncu.log 2>&1
tail -2 gpurun_out/r02_pseudo_ncu.log
cat gpurun_out/round2_first_call.txt
