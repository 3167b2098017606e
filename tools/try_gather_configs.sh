#!/bin/bash
# Rebuild with different output-tile shapes of the fused gather (rows per thread, staging KB, thread rows, tile width) and print phase times (GPU box only).
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r rp kb ty tw <<< "$cfg"
  rm -f cmda_b200/csrc/build/voxel_factored.o
  make -C cmda_b200/csrc -j8 EXTRA="-DCMDA_OUT_ROWS_PER_THREAD=$rp -DCMDA_STAGE_KB=$kb -DCMDA_OUT_TY=${ty:-8} -DCMDA_OUT_W=${tw:-64}" > /dev/null 2>&1 || { echo "$cfg build failed"; continue; }
  for b in 5 1; do
    python bench.py --steps 20 --warmup 3 --bins $b --no-cpu-baseline --no-pseudo --no-variants 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$cfg B=$b', round(d['value']), round(d['ms_per_step'],3), {k[:14]:round(v,3) for k,v in d['roofline']['phase_ms'].items()})"
  done
done
rm -f cmda_b200/csrc/build/voxel_factored.o
make -C cmda_b200/csrc -j8 > /dev/null 2>&1
