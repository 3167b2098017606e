#!/bin/bash
# Round-2 evidence of the current build on one B200 (run under gpurun; outputs in gpurun_out/, summaries are copied to
# profiles/ afterwards): GPU test suite, smoke, bench lines of both arms, the ncu launch list of the bench command, one
# `--set full` capture of the voxel kernels (+ DRAM traffic JSON labelled with the commit) and of the pseudo-event kernels.
cd "$(dirname "$0")/.."
SHA=${1:-unknown}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_bench_b5.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pseudo --no-variants --no-c4 > gpurun_out/r02_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_voxel_b5 \
    -k regex:"sensor_accumulate|rectify_gather|norm_apply|rectify_index|stencil_build|out_tile_box|regroup|fallback" -c 20 \
    python tools/profile_step.py --bins 5 --mode auto --steps 2 > gpurun_out/r02_ncu_voxel.log 2>&1
python tools/ncu_traffic_json.py gpurun_out/r02_voxel_b5.ncu-rep "$SHA" gpurun_out/ncu_traffic.json > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_voxel_b1 \
    -k regex:"band_partition|band_accumulate|rectify_gather|norm_apply" -c 8 \
    python tools/profile_step.py --bins 1 --mode auto --steps 2 > gpurun_out/r02_ncu_voxel_b1.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_pseudo_final \
    -k regex:"pair_|isr_" -c 6 python tools/profile_pseudo.py > gpurun_out/r02_ncu_pseudo.log 2>&1
head -c 400 gpurun_out/r02_bench_n1.json; echo; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
