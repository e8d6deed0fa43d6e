#!/bin/bash
# round 2, GPU call 6 (1 GPU): fused TD+train kernel, env one-wave rule; tests, bench, variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c6_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c6_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?" >> gpurun_out/c6_bench.err
rm -f gpurun_out/c6_variants.txt
for v in "--algo ddqn" "--envs 16384" "--envs 65536 --algo ddqn" "--envs 16384 --net vanet2 --algo dueling" "--net vanet2 --algo dueling"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.3 $v 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(sys.argv[1:], round(d['value']/1e6,2),'M steps/s',round(d['ms_per_step']*1e3,2),'us/iter',{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" $v >> gpurun_out/c6_variants.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 100 --csv --log-file gpurun_out/c6_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c6_ncu_launch.log 2>&1
tail -6 gpurun_out/c6_pytest.txt; head -c 300 gpurun_out/c6_bench.json; tail -2 gpurun_out/c6_bench.err; cat gpurun_out/c6_variants.txt
