#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q -k "pipeline or p3 or packed" 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pseudo --no-variants --no-c4 2> gpurun_out/r02_call25.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'])
print(json.dumps(d['e2e'])[:600])
for k,v in d['e2e_other_wires'].items(): print(k, v['value'], v['ms_per_step'], v['matches_device_path'])
"
tail -3 gpurun_out/r02_call25.err
} > gpurun_out/r02_call25.txt 2>&1
cat gpurun_out/r02_call25.txt
