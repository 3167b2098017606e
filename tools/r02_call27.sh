#!/bin/bash
mkdir -p gpurun_out
{
timeout 120 python tools/phase_times.py --bins 5
timeout 120 python tools/phase_times.py --bins 1 --store soa
timeout 300 python tools/small_call_times.py
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
} > gpurun_out/r02_call27.txt 2>&1
cat gpurun_out/r02_call27.txt
