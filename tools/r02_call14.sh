#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "events_vg or c1_ or c2_ or hot_pixel or fused or packed or pipeline or dsec" > gpurun_out/r02_pytest_gpu_j.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu_j.txt
for b in 5 1; do
timeout 300 python bench.py --steps 20 --warmup 5 --bins $b --no-cpu-baseline --no-variants --no-pseudo --no-c4 2>gpurun_out/r02_j.err | python -c "
import json,sys;d=json.load(sys.stdin);print('B=$b', d['resolved_mode'], round(d['ms_per_step'],3), {k[:18]:round(v,3) for k,v in d['roofline']['phase_ms'].items()}, round(d['e2e']['value']), d['e2e']['host_link_probe'])"
done
