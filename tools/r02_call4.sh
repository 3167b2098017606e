#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_c.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_c.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-variants --no-pseudo 2>gpurun_out/r02_bench_c.err | tee gpurun_out/r02_bench_c.json | python -c "
import json,sys;d=json.load(sys.stdin);print(d['resolved_mode'], round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['phase_ms'].items()}); print('e2e', d['e2e']); print(d['e2e_other_wires']); print(d['cpu_baseline']); print(d['cpu_port'])"
tail -3 gpurun_out/r02_bench_c.err
