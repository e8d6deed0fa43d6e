#!/bin/bash
# round 2, GPU call 8 (2 GPUs): stacked 3xTF32 + block-wise data-parallel optimiser kernel: full tests, 1- and 2-GPU bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c8_bench_1gpu.json 2> gpurun_out/c8_bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c8_bench_2gpu.json 2> gpurun_out/c8_bench_2gpu.err; echo "rc=$?" >> gpurun_out/c8_bench_2gpu.err
UAVRL_DP_TWO_KERNELS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-configs > gpurun_out/c8_bench_2gpu_pair.json 2> gpurun_out/c8_bench_2gpu_pair.err
UAVRL_TC_TRACE=1 timeout 200 python tools/tc_trace.py 2>&1 | grep "tc_trace" | tail -2 > gpurun_out/c8_trace.txt
tail -5 gpurun_out/c8_pytest.txt
for f in c8_bench_1gpu c8_bench_2gpu c8_bench_2gpu_pair; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})
except Exception as e: print('$f', 'ERR', e)
"; done; tail -3 gpurun_out/c8_bench_2gpu.err; cat gpurun_out/c8_trace.txt
