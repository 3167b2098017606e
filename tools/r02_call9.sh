#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "isr or image_change or pseudo or pair or source_img or mixed" > gpurun_out/r02_pytest_gpu_h.txt 2>&1
tail -3 gpurun_out/r02_pytest_gpu_h.txt
for lib in "" cmda_b200/variants/lib_notab.so; do
CMDA_B200_LIB=$lib python - <<'PY'
import sys, json, os
sys.path.insert(0, '.')
import torch, bench
dev = torch.device('cuda:0')
r = bench.pseudo_events_leg(dev, 6553.6)
print(os.environ.get('CMDA_B200_LIB') or 'tables', {k: (round(v['ms_per_batch'], 4), round(v['frac_of_hbm_peak'], 3)) for k, v in r.items() if isinstance(v, dict)})
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo_tab3 \
    -k regex:"pair_|isr_" -c 7 python tools/profile_pseudo.py > gpurun_out/r02_pseudo_tab3_ncu.log 2>&1
tail -2 gpurun_out/r02_pseudo_tab3_ncu.log
