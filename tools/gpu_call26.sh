#!/bin/bash
# round 2, GPU call 26 (2 GPUs): diagnose the 2-GPU slowdown seen in call 25
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
nvidia-smi --query-compute-apps=pid,used_memory --format=csv > gpurun_out/c26_apps0.txt 2>&1
UAVRL_DP_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $B > gpurun_out/c26_a.json 2> gpurun_out/c26_a.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --dp nccl $B > gpurun_out/c26_b.json 2> gpurun_out/c26_b.err
UAVRL_DP_TWO_KERNELS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 $B > gpurun_out/c26_c.json 2> gpurun_out/c26_c.err
UAVRL_DP_TRACE=1 timeout 300 python bench.py --gpus 1 --dp-self 1 $B > gpurun_out/c26_d.json 2> gpurun_out/c26_d.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --pdl 0 $B > gpurun_out/c26_e.json 2> gpurun_out/c26_e.err
for f in a b c d e; do python -c "
import json
d=json.loads(open('gpurun_out/c26_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'blocks', d.get('repeats'), round(d.get('block_ms_min',0),2), round(d.get('block_ms_median',0),2), round(d.get('block_ms_max',0),2))"; grep -h dp_trace gpurun_out/c26_$f.err | tail -2; done
