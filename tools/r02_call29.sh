#!/bin/bash
mkdir -p gpurun_out
timeout 800 python tools/gpu_fuzz.py ${1:-100} ${2:-480} > gpurun_out/r02_gpu_fuzz.txt 2>&1
tail -n 8 gpurun_out/r02_gpu_fuzz.txt
