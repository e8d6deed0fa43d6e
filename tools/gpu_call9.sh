#!/bin/bash
# round 2, GPU call 9 (2 GPUs): stacked epilogue with overlapped TMEM loads; data-parallel phase trace (world 1 and 2)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_train_gpu.py tests/test_learner_gpu.py tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c9_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c9_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c9_bench_1gpu.json 2> gpurun_out/c9_bench_1gpu.err
UAVRL_DP_TRACE=1 timeout 300 python bench.py --gpus 1 --dp-self 1 $B > gpurun_out/c9_bench_1gpu_dpself.json 2> gpurun_out/c9_bench_1gpu_dpself.err
UAVRL_DP_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/c9_bench_2gpu.json 2> gpurun_out/c9_bench_2gpu.err
UAVRL_DP_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/c9_bench_2gpu_k200.json 2> gpurun_out/c9_bench_2gpu_k200.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --dp nccl $B > gpurun_out/c9_bench_2gpu_nccl.json 2> gpurun_out/c9_bench_2gpu_nccl.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "tc_trace" | tail -2 > gpurun_out/c9_trace.txt
tail -5 gpurun_out/c9_pytest.txt
for f in c9_bench_1gpu c9_bench_1gpu_dpself c9_bench_2gpu c9_bench_2gpu_k200 c9_bench_2gpu_nccl; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d.get('kernels',{}).items()})
except Exception as e: print('$f', 'ERR', e)
"; grep -h "dp_trace" gpurun_out/$f.err | tail -4; done; cat gpurun_out/c9_trace.txt
