#!/bin/bash
# Round-end evidence of the current build, one GPU (run under gpurun): bench lines, the ncu launch list of the bench
# command, one `--set full` capture of the FACTORED kernels, smoke() and the GPU test suite.  Outputs in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err
python bench.py --impl reference > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pseudo --no-variants > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r01_factored_b5 \
    -k regex:"sensor_accumulate|rectify_gather|norm_apply|rectify_index|stencil_build|out_tile_box|regroup" -c 16 \
    python tools/profile_step.py --bins 5 --mode factored --steps 1 > gpurun_out/ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1
tail -1 gpurun_out/smoke.log; tail -1 gpurun_out/pytest_gpu.txt; head -c 600 gpurun_out/bench_r01_n1.json; echo; cat gpurun_out/bench_r01_ref.json | head -c 600
