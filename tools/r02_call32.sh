#!/bin/bash
mkdir -p gpurun_out
python tools/profile_p3.py > gpurun_out/r02_p3_unpack.txt 2>&1
ncu --set full --clock-control none -f -o gpurun_out/r02_p3 -k regex:p3_unpack -c 1 python tools/profile_p3.py > /dev/null 2>&1
{ echo "# ncu --set full --clock-control none, tools/profile_p3.py (10 M events)"; python tools/ncu_raw_summary.py gpurun_out/r02_p3.ncu-rep; } >> gpurun_out/r02_p3_unpack.txt 2>&1
rm -f gpurun_out/r02_p3.ncu-rep
cat gpurun_out/r02_p3_unpack.txt
