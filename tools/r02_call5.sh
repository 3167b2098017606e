#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "banded or c1_window or packed_store or hot_pixel" > gpurun_out/r02_pytest_gpu_d.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_d.txt
for m in banded banded2; do
  for b in 5 1; do
    timeout 120 python bench.py --steps 20 --warmup 3 --bins $b --mode $m --no-cpu-baseline --no-variants --no-pseudo 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$m B=$b', round(d['ms_per_step'],3), [round(v,3) for v in d['roofline']['phase_ms'].values()], d['e2e']['matches_device_path'])"
  done
done
