#!/bin/bash
# round 2, GPU call 31 (4 GPUs): final code, the driver's command at N = 4 (+ the exchange's phase trace)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c31_bench_4gpu.json 2> gpurun_out/c31_bench_4gpu.err; echo "rc=$?" >> gpurun_out/c31_bench_4gpu.err
UAVRL_DP_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-configs > gpurun_out/c31_bench_4gpu_tr.json 2> gpurun_out/c31_bench_4gpu_tr.err
for f in c31_bench_4gpu c31_bench_4gpu_tr; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), 'blocks', d.get('repeats'), round(d.get('block_ms_min',0),2), round(d.get('block_ms_median',0),2), round(d.get('block_ms_max',0),2))"; grep -h dp_trace gpurun_out/$f.err | tail -4; done
