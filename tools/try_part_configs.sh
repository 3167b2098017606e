#!/bin/bash
# Rebuild libcmda_b200.so with different partition-kernel shapes and print the phase times (GPU box only).
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  IFS=, read -r th gr mb <<< "$cfg"
  rm -f cmda_b200/csrc/build/voxel_tiled.o
  make -C cmda_b200/csrc -j8 EXTRA="-DCMDA_PART_THREADS=$th -DCMDA_PART_GROUPS=$gr -DCMDA_PART_MINBLOCKS=$mb" > /dev/null 2>&1 || { echo "$cfg build failed"; continue; }
  for b in 5 1; do
    python bench.py --steps 10 --warmup 3 --bins $b --mode tiled --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$cfg B=$b', round(d['value']), round(d['ms_per_step'],3), {k[:14]:round(v,3) for k,v in d['roofline']['phase_ms'].items()})"
  done
done
rm -f cmda_b200/csrc/build/voxel_tiled.o
make -C cmda_b200/csrc -j8 > /dev/null 2>&1
