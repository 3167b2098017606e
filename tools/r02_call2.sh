#!/bin/bash
# round 2, GPU call 2: the whole GPU suite with the new full-size parity tests + a short bench (sketch overhead, AUTO at B=1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_a.txt 2>&1
tail -5 gpurun_out/r02_pytest_gpu_a.txt
for b in 5 1; do
  timeout 300 python bench.py --steps 20 --warmup 3 --bins $b --no-cpu-baseline --no-variants --no-pseudo 2>gpurun_out/r02_bench_a_b$b.err | tee gpurun_out/r02_bench_a_b$b.json | python -c "
import json,sys;d=json.load(sys.stdin);print('B=$b', d['resolved_mode'], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['phase_ms'].items()}, d['e2e']['matches_device_path'], round(d['e2e']['ms_per_step'],2))"
done
