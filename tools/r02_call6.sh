#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "banded or isr or image_change or pseudo or pair or source_img or mixed" > gpurun_out/r02_pytest_gpu_e.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_e.txt
for m in banded banded2; do
  for b in 5 1; do
    timeout 120 python bench.py --steps 20 --warmup 3 --bins $b --mode $m --no-cpu-baseline --no-variants --no-pseudo 2>gpurun_out/r02_e_$m$b.err | python -c "
import json,sys;d=json.load(sys.stdin);print('$m B=$b', round(d['ms_per_step'],3), [round(v,3) for v in d['roofline']['phase_ms'].values()], d['e2e']['matches_device_path'])"
  done
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-variants 2>gpurun_out/r02_e_pseudo.err | python -c "
import json,sys;d=json.load(sys.stdin);print(json.dumps(d['pseudo_events'], indent=1)); print(json.dumps(d['train_step_input_path']))"
