#!/bin/bash
# round 2, GPU call 16 (1 GPU): env statistics off the chain, head stage marks, fused act+env variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_train_gpu.py tests/test_learner_gpu.py tests/test_env_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c16_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c16_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c16_bench_1gpu.json 2> gpurun_out/c16_bench_1gpu.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -9 > gpurun_out/c16_trace.txt
UAVRL_ENV_TRACE=1 timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs --min-seconds 0.05 2>&1 | grep "env_trace" | tail -3 >> gpurun_out/c16_trace.txt
tail -4 gpurun_out/c16_pytest.txt
for f in c16_bench_1gpu; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d.get('kernels',{}).items()})
except Exception as e: print('$f', 'ERR', e)
"; done; cat gpurun_out/c16_trace.txt
timeout 300 python bench.py --gpus 1 --fuse 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/c16_bench_fuse1.json 2> gpurun_out/c16_bench_fuse1.err
python -c "
import json
d=json.loads(open('gpurun_out/c16_bench_fuse1.json').read().strip().splitlines()[-1]); print('fuse1', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter')"
