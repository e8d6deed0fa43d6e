#!/bin/bash
# round 2, GPU call 2: tests after the MN-major dW kernel / fused optimiser tail / push all-reduce / env chain changes, bench, launch list, traces
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/c2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?" >> gpurun_out/c2_bench.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py > gpurun_out/c2_trace.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 540 -c 100 --csv --log-file gpurun_out/c2_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c2_ncu_launch.log 2>&1
for v in "--envs 16384" "--envs 65536 --algo ddqn"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.3 $v 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(sys.argv[1:], round(d['value']/1e6,2),'M steps/s',round(d['ms_per_step']*1e3,2),'us/iter',{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" $v >> gpurun_out/c2_variants.txt 2>&1
done
tail -5 gpurun_out/c2_pytest.txt; head -c 400 gpurun_out/c2_bench.json; tail -2 gpurun_out/c2_bench.err; cat gpurun_out/c2_variants.txt
