#!/bin/bash
# GPU box: phase times of the BANDED stage A for the shipped build (B = 5 and 1) and every variant in
# cmda_b200/variants/ (B = $VBINS, default 5).  Phases: memset | plans | partition | accumulate | gather | norm.
cd "$(dirname "$0")/.."
run() {  # lib, bins
  CMDA_B200_LIB=$PWD/$1 timeout 90 python bench.py --steps 20 --warmup 3 --bins $2 --mode ${MODE:-banded} --no-cpu-baseline --no-variants --no-pseudo 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$(basename $1) B=$2', round(d['ms_per_step'],3), [round(v,3) for v in d['roofline']['phase_ms'].values()], d['e2e']['matches_device_path'])"
}
for b in 5 1; do run cmda_b200/libcmda_b200.so $b; done
for lib in cmda_b200/variants/lib_*.so; do
  [ -f "$lib" ] || continue
  for b in ${VBINS:-5}; do run $lib $b; done
done
