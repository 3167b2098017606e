#!/bin/bash
# scaling evidence of the final build at N = all GPUs of the box: both wires, C4 strong scaling, topology
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 \
    --no-variants --no-pseudo > gpurun_out/r02_bench_final_n$N.json 2> gpurun_out/r02_bench_final_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_final_n$N.json'))
o = d['e2e_other_wires']
print('N=$N value', round(d['value']), 'e2e_p3', round(d['e2e']['value']), 'p4', round(o['p4']['value']), 'soa', round(o['soa']['value']), 'resident', round(o['resident']['value']),
      'h2d/gpu', round(d['e2e']['h2d_GBps_per_gpu'], 1), 'probe', d["e2e"]["host_link_probe"], 'c4', round(d['c4_strong_scaling']['Mevents_per_s']))
PY
tail -3 gpurun_out/r02_bench_final_n$N.err
