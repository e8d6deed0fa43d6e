#!/bin/bash
# round 2, GPU call 21 (1 GPU): rows resolved before the PDL wait, lean dW epilogue stores, fence-free fused dW + Adam tail
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_train_gpu.py tests/test_learner_gpu.py tests/test_per_gpu.py tests/test_multigpu_gpu.py tests/test_plugins_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c21_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c21_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c21_bench_1gpu.json 2> gpurun_out/c21_bench_1gpu.err
timeout 300 python bench.py --gpus 1 --fuse-dw 1 $B > gpurun_out/c21_bench_fusedw.json 2> gpurun_out/c21_bench_fusedw.err
timeout 300 python bench.py --gpus 1 --algo ddqn $B > gpurun_out/c21_bench_ddqn.json 2> gpurun_out/c21_bench_ddqn.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -3 > gpurun_out/c21_trace.txt
tail -4 gpurun_out/c21_pytest.txt
for f in c21_bench_1gpu c21_bench_fusedw c21_bench_ddqn; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; done
cat gpurun_out/c21_trace.txt
