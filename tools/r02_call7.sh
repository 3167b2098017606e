#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for m in banded banded2; do
  for b in 5 1; do
    timeout 120 python bench.py --steps 20 --warmup 3 --bins $b --mode $m --no-cpu-baseline --no-variants --no-pseudo 2>gpurun_out/r02_f_$m-$b.err | python -c "
import json,sys;d=json.load(sys.stdin);print('$m B=$b', round(d['ms_per_step'],3), [round(v,3) for v in d['roofline']['phase_ms'].values()], d['e2e']['matches_device_path'])"
  done
done > gpurun_out/r02_f_banded.txt 2>&1
cat gpurun_out/r02_f_banded.txt
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo_tab \
    -k regex:"pair_|isr_" -c 14 python tools/profile_pseudo.py > gpurun_out/r02_pseudo_tab_ncu.log 2>&1
tail -2 gpurun_out/r02_pseudo_tab_ncu.log
