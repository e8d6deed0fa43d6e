#!/bin/bash
# round 2, GPU call 13 (1 GPU): launch list + full source-level ncu capture of the five loop kernels (1 M-transition ring, the timed configuration)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 100 --csv --log-file gpurun_out/c13_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c13_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|env_kernel|reduce_adam' -s 536 -c 10 -o gpurun_out/c13_prof \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c13_ncu_full.log 2>&1
tail -3 gpurun_out/c13_ncu_full.log; ls -la gpurun_out/c13_prof.ncu-rep; tail -12 gpurun_out/c13_launches.csv | cut -c1-160
