#!/bin/bash
mkdir -p gpurun_out
bash tools/r02_call23.sh > /dev/null 2>&1
tail -n 12 gpurun_out/r02_sanitizer.txt
timeout 200 python tools/gpu_fuzz.py 20000 150 2>&1 | tail -n 3
