#!/bin/bash
# round 2, GPU call 22 (1 GPU): image staging in pieces, next-state rows gathered before the wait, Adam state prefetched before the wait
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_train_gpu.py tests/test_learner_gpu.py tests/test_per_gpu.py tests/test_multigpu_gpu.py tests/test_plugins_gpu.py tests/test_sac_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c22_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c22_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c22_bench_1gpu.json 2> gpurun_out/c22_bench_1gpu.err
timeout 300 python bench.py --gpus 1 --algo ddqn --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/c22_bench_ddqn.json 2> gpurun_out/c22_bench_ddqn.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -4 > gpurun_out/c22_trace.txt
tail -4 gpurun_out/c22_pytest.txt
for f in c22_bench_1gpu c22_bench_ddqn; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, {k:round(v['value']/1e6,1) for k,v in d.get('configs',{}).items()})"; done
cat gpurun_out/c22_trace.txt
