#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
tail -3 gpurun_out/r02_bench_full.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench_full.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step'])
print('e2e', d['e2e']['value'], {k: round(v['value']) for k, v in d['e2e_other_wires'].items()})
print('c4', d['c4_strong_scaling'])
print('cpu', d['cpu_baseline']['value'], d['cpu_port']['value'])
print('pseudo', {k: (v.get('ms_per_batch'), v.get('frac_of_hbm_peak'), v.get('value')) for k, v in d['pseudo_events'].items() if isinstance(v, dict)})
print('c5', d['train_step_input_path'])
print('variants', {k: round(v['ms_per_step'], 3) for k, v in d['variants'].items()})
print('exp', d['experimental'])
PY
