#!/bin/bash
# round 2, GPU call 11 (2 GPUs): fence-free {epoch:value} word exchange in the data-parallel kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c11_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c11_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
UAVRL_DP_TRACE=1 timeout 300 python bench.py --gpus 1 --dp-self 1 $B > gpurun_out/c11_bench_1gpu_dpself.json 2> gpurun_out/c11_bench_1gpu_dpself.err
timeout 300 python bench.py --gpus 1 --dp-self 1 $B > gpurun_out/c11_bench_1gpu_dpself_nt.json 2> gpurun_out/c11_bench_1gpu_dpself_nt.err
UAVRL_DP_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/c11_bench_2gpu_tr.json 2> gpurun_out/c11_bench_2gpu_tr.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 $B > gpurun_out/c11_bench_2gpu.json 2> gpurun_out/c11_bench_2gpu.err
UAVRL_DP_TWO_KERNELS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $B > gpurun_out/c11_bench_2gpu_pair.json 2> gpurun_out/c11_bench_2gpu_pair.err
tail -3 gpurun_out/c11_pytest.txt
for f in c11_bench_1gpu_dpself c11_bench_1gpu_dpself_nt c11_bench_2gpu_tr c11_bench_2gpu c11_bench_2gpu_pair; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d.get('kernels',{}).items()})
except Exception as e: print('$f', 'ERR', e)
"; grep -h "dp_trace" gpurun_out/$f.err | tail -4; done
