#!/bin/bash
# First GPU call of round 2 (about 90 s on one B200): everything that was prepared without a GPU at the end of round 1.
#   1. BANDED2 on hardware for the first time: bit-identity against FACTORED (also with polarity bytes beyond {0, 1})
#      and ms per step next to FACTORED and BANDED on the C2 workload, B = 5 and 1   (tools/banded2_check.py)
#   2. phase times (partition | accumulate) of BANDED and BANDED2                     (bench.py --mode ...)
#   3. the windows of a step split across two streams, FACTORED | BANDED              (tools/split_probe.py)
# Outputs in gpurun_out/round2_first_call.txt.   Use: gpurun --timeout 300 -- tools/round2_first_call.sh
# Optional, BEFORE the call (here, on the CPU; the .so files travel): kernel-shape variants that the sweep at the end picks up
#   tools/build_variant.sh b2p512  "-DCMDA_BAND2_PART_THREADS=512 -DCMDA_BAND2_PART_GROUPS=2"   # BANDED2 partition at 64 registers
#   tools/build_variant.sh acc1024 "-DCMDA_BAND_ACC_THREADS=1024"                               # 64 resident warps in the accumulate pass
#   tools/build_variant.sh u8      "-DCMDA_BAND_V2_UNROLL=8"                                    # 256-record rounds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== banded2_check"; timeout 120 python tools/banded2_check.py 2>&1 | tail -3
  for m in banded banded2; do
    for b in 5 1; do
      timeout 90 python bench.py --steps 20 --warmup 3 --bins $b --mode $m --no-cpu-baseline --no-variants --no-pseudo 2>/dev/null | python -c "
import json,sys;d=json.load(sys.stdin);print('$m B=$b', round(d['ms_per_step'],3), [round(v,3) for v in d['roofline']['phase_ms'].values()], d['e2e']['matches_device_path'])"
    done
  done
  echo "== split_probe"; timeout 90 python tools/split_probe.py --bins 5 2>&1 | tail -8; timeout 90 python tools/split_probe.py --bins 5 --second banded2 --splits 4,6,8,10 2>&1 | tail -5
  if ls cmda_b200/variants/lib_*.so > /dev/null 2>&1; then echo "== variants (banded2)"; MODE=banded2 VBINS="5 1" tools/banded_sweep.sh 2>&1 | grep lib_; fi
} | tee gpurun_out/round2_first_call.txt
