#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "exact_mode_bit_identical_at_5m or c4_window or hot_pixel or auto_mode or banded2_identical or denorm_to_gray or mixed_image" > gpurun_out/r02_pytest_gpu_b.txt 2>&1
tail -5 gpurun_out/r02_pytest_gpu_b.txt
for b in 5; do
  timeout 300 python bench.py --steps 20 --warmup 3 --bins $b --no-cpu-baseline --no-variants --no-pseudo 2>gpurun_out/r02_bench_b_b$b.err | tee gpurun_out/r02_bench_b_b$b.json | python -c "
import json,sys;d=json.load(sys.stdin);print('B=$b', d['resolved_mode'], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['phase_ms'].items()}, d['e2e']['matches_device_path'], round(d['e2e']['ms_per_step'],2))"
done
