#!/bin/bash
# round 2, GPU call 29 (1 GPU): final launch list + full ncu capture (DRAM traffic per launch) on the timed configuration (1 M ring); driver-style bench + reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 100 --csv --log-file gpurun_out/c29_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c29_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|env_kernel|reduce_adam' -s 536 -c 10 -o gpurun_out/c29_prof \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c29_ncu_full.log 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c29_bench_1gpu.json 2> gpurun_out/c29_bench_1gpu.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/c29_bench_reference.json 2> gpurun_out/c29_bench_reference.err
ls -la gpurun_out/c29_*; head -c 600 gpurun_out/c29_bench_reference.json
