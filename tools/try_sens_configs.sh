#!/bin/bash
# Rebuild with different stage-A kernel shapes (threads,groups,minblocks,prefetch) and print phase times (GPU box only).
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r th gr mb pf <<< "$cfg"
  rm -f cmda_b200/csrc/build/voxel_factored.o
  make -C cmda_b200/csrc -j8 EXTRA="-DCMDA_SENS_THREADS=$th -DCMDA_SENS_GROUPS=$gr -DCMDA_SENS_MINBLOCKS=$mb -DCMDA_SENS_PREFETCH=$pf" > /dev/null 2>&1 || { echo "$cfg build failed"; continue; }
  for b in 5 1; do
    python bench.py --steps 10 --warmup 3 --bins $b --mode factored --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$cfg B=$b', round(d['value']), round(d['ms_per_step'],3), {k[:14]:round(v,3) for k,v in d['roofline']['phase_ms'].items()})"
  done
done
rm -f cmda_b200/csrc/build/voxel_factored.o
make -C cmda_b200/csrc -j8 > /dev/null 2>&1
