#!/bin/bash
# 8-GPU scaling evidence: bench arm + reference arm at N = 8 (and N = 4), topology
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for n in $N 4; do
  [ $n -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 20 --warmup 5 \
      --no-variants --no-pseudo > gpurun_out/r02_bench_scale_n$n.json 2> gpurun_out/r02_bench_scale_n$n.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_scale_n$n.json'))
print('N=$n value', round(d['value']), 'e2e', round(d['e2e']['value']), 'soa', round(d['e2e_other_wires']['soa']['value']), 'resident', round(d['e2e_other_wires']['resident']['value']),
      'h2d/gpu', round(d['e2e']['h2d_GBps_per_gpu'], 1), 'probe', d["e2e"]["host_link_probe"], 'c4', round(d['c4_strong_scaling']['Mevents_per_s']))
PY
done
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1; nproc >> gpurun_out/r02_topo_n8.txt; free -g >> gpurun_out/r02_topo_n8.txt
