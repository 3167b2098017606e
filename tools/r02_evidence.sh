#!/bin/bash
# Round-2 evidence of the current build on one B200 (run under gpurun; outputs in gpurun_out/, summaries are copied to
# profiles/ afterwards).  `tools/r02_evidence.sh <sha> run`: GPU test suite, smoke, bench lines of both arms, the ncu
# launch list of the bench command.  `tools/r02_evidence.sh <sha> ncu`: `--set full` captures of the voxel kernels
# (+ DRAM traffic JSON labelled with the commit) and of the pseudo-event kernels, summarised ON the box -- gpurun brings
# back at most 64 MiB, so only the B = 5 report itself travels.
cd "$(dirname "$0")/.."
SHA=${1:-unknown}
WHAT=${2:-run}
mkdir -p gpurun_out
if [ "$WHAT" = run ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_b5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pseudo --no-variants --no-c4 > gpurun_out/r02_bench_under_ncu.log 2>&1
head -c 400 gpurun_out/r02_bench_n1.json; echo; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
else
HDR="# ncu --set full --clock-control none (cold caches, serialised), commit $SHA"
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_voxel_b5 \
    -k regex:"sensor_accumulate|rectify_gather|norm_apply|rectify_index|stencil_build|out_tile_box|regroup|fallback" -c 20 \
    python tools/profile_step.py --bins 5 --mode auto --steps 2 > gpurun_out/r02_ncu_voxel.log 2>&1
python tools/ncu_traffic_json.py gpurun_out/r02_voxel_b5.ncu-rep "$SHA" gpurun_out/ncu_traffic.json > /dev/null 2>&1
{ echo "$HDR, tools/profile_step.py --bins 5 --mode auto --store p4 (C2: 16 x 5 M events, packed store, plans rebuilt per step)"
  python tools/ncu_raw_summary.py gpurun_out/r02_voxel_b5.ncu-rep
  for k in sensor_accumulate rectify_gather norm_apply; do echo "== $k: opcode mix and hottest SASS lines"; python tools/ncu_source_summary.py gpurun_out/r02_voxel_b5.ncu-rep $k 14; done
} > gpurun_out/r02_ncu_voxel_b5_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_voxel_b1 \
    -k regex:"band_partition|band_accumulate|rectify_gather|norm_apply" -c 8 \
    python tools/profile_step.py --bins 1 --mode auto --steps 2 > gpurun_out/r02_ncu_voxel_b1.log 2>&1
{ echo "$HDR, tools/profile_step.py --bins 1 --mode auto --store p4 (C2 at the shipped events_bins = 1: AUTO takes the BANDED cut)"
  python tools/ncu_raw_summary.py gpurun_out/r02_voxel_b1.ncu-rep
  for k in band_partition3 band_accumulate; do echo "== $k: opcode mix and hottest SASS lines"; python tools/ncu_source_summary.py gpurun_out/r02_voxel_b1.ncu-rep $k 14; done
} > gpurun_out/r02_ncu_voxel_b1_summary.txt 2>&1
rm -f gpurun_out/r02_voxel_b1.ncu-rep
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo_final \
    -k regex:"pair_|isr_" -c 6 python tools/profile_pseudo.py > gpurun_out/r02_ncu_pseudo.log 2>&1
{ echo "$HDR, tools/profile_pseudo.py (C3: 32 x 2048x1024; frame pair f32+u8, frame pair u8 only (table pass), shift pair)"
  python tools/ncu_raw_summary.py gpurun_out/r02_pseudo_final.ncu-rep
} > gpurun_out/r02_ncu_pseudo_summary.txt 2>&1
rm -f gpurun_out/r02_pseudo_final.ncu-rep
ls -la gpurun_out | head -30
fi
