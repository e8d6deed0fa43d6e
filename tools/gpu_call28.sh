#!/bin/bash
# round 2, GPU call 28 (1 GPU): compile-time specialised tensor-core kernels (mode / stacking / dueling / TD passes / fused tail)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c28_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c28_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c28_bench_1gpu.json 2> gpurun_out/c28_bench_1gpu.err
timeout 300 python bench.py --gpus 1 --algo ddqn $B > gpurun_out/c28_bench_ddqn.json 2> gpurun_out/c28_bench_ddqn.err
timeout 300 python bench.py --gpus 1 --envs 16384 --net vanet2 --algo dueling $B > gpurun_out/c28_bench_16k.json 2> gpurun_out/c28_bench_16k.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -3 > gpurun_out/c28_trace.txt
tail -3 gpurun_out/c28_pytest.txt
for f in c28_bench_1gpu c28_bench_ddqn c28_bench_16k; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; done
cat gpurun_out/c28_trace.txt
