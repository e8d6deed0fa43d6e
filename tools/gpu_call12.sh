#!/bin/bash
# round 2, GPU call 12 (8 GPUs): the driver's scaling commands at N = 8 and N = 4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c12_topo.txt 2>&1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c12_bench_8gpu.json 2> gpurun_out/c12_bench_8gpu.err; echo "rc=$?" >> gpurun_out/c12_bench_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-configs > gpurun_out/c12_bench_4gpu.json 2> gpurun_out/c12_bench_4gpu.err; echo "rc=$?" >> gpurun_out/c12_bench_4gpu.err
UAVRL_DP_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-configs > gpurun_out/c12_bench_8gpu_tr.json 2> gpurun_out/c12_bench_8gpu_tr.err; echo "rc=$?" >> gpurun_out/c12_bench_8gpu_tr.err
for f in c12_bench_8gpu c12_bench_4gpu c12_bench_8gpu_tr; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['value']/1e6,1) for k,v in d.get('configs',{}).items()})
except Exception as e: print('$f', 'ERR', e)
"; grep -h "dp_trace" gpurun_out/$f.err | tail -8; tail -2 gpurun_out/$f.err; done
