#!/bin/bash
# round 2, GPU call 1: parity tests (incl. the new large-variant ones), layout probe, bench (both arms), ncu launch list + full capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt; python -c "import os;print(len(os.sched_getaffinity(0)))" >> gpurun_out/c1_nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/c1_nproc.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.txt
timeout 120 ./tools/umma_mn_probe > gpurun_out/c1_mn_probe.txt 2>&1; echo "probe rc=$?" >> gpurun_out/c1_mn_probe.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err; echo "ref rc=$?" >> gpurun_out/c1_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 150 --csv --log-file gpurun_out/c1_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c1_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|env_kernel|reduce_adam' -s 530 -c 12 -o gpurun_out/c1_prof \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c1_ncu_full.log 2>&1
ls -la gpurun_out > gpurun_out/c1_ls.txt
tail -5 gpurun_out/c1_pytest.txt; tail -3 gpurun_out/c1_mn_probe.txt; head -c 600 gpurun_out/c1_bench.json; tail -2 gpurun_out/c1_bench.err
