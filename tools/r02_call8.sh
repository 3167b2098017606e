#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "isr or image_change or pseudo or pair or source_img or mixed" > gpurun_out/r02_pytest_gpu_g.txt 2>&1
tail -3 gpurun_out/r02_pytest_gpu_g.txt
python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch, bench
dev = torch.device('cuda:0')
print(json.dumps(bench.pseudo_events_leg(dev, 6553.6), indent=1))
PY
timeout 300 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo_tab2 \
    -k regex:"pair_|isr_" -c 7 python tools/profile_pseudo.py > gpurun_out/r02_pseudo_tab2_ncu.log 2>&1
tail -2 gpurun_out/r02_pseudo_tab2_ncu.log
