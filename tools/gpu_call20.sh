#!/bin/bash
# round 2, GPU call 20 (1 GPU): control-warp prologues; full GPU test suite, bench, stage traces, SAC loop
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c20_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c20_bench_1gpu.json 2> gpurun_out/c20_bench_1gpu.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -5 > gpurun_out/c20_trace.txt
timeout 300 python tools/sac_probe.py 16384 20 1048576 > gpurun_out/c20_sac.txt 2>&1
tail -4 gpurun_out/c20_pytest.txt; grep sac_probe gpurun_out/c20_sac.txt
python -c "
import json
d=json.loads(open('gpurun_out/c20_bench_1gpu.json').read().strip().splitlines()[-1]); print(round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"
cat gpurun_out/c20_trace.txt
