#!/bin/bash
# round 2, GPU call 3: full tests, bench, launch list, source-level profile of env / dW / forward kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?" >> gpurun_out/c3_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 100 --csv --log-file gpurun_out/c3_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c3_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|env_kernel|reduce_adam' -s 536 -c 6 -o gpurun_out/c3_prof \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c3_ncu_full.log 2>&1
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -3 > gpurun_out/c3_trace.txt
tail -8 gpurun_out/c3_pytest.txt; head -c 300 gpurun_out/c3_bench.json; tail -2 gpurun_out/c3_bench.err
