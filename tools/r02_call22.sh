#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/overlap_probe.py --bins 5
timeout 300 python tools/overlap_probe.py --bins 1
} > gpurun_out/r02_overlap_probe.txt 2>&1
cat gpurun_out/r02_overlap_probe.txt
