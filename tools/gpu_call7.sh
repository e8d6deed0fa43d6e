#!/bin/bash
# round 2, GPU call 7 (2 GPUs): two-rank data-parallel tests (push all-reduce fused with Adam) and the 2-GPU bench, both --dp modes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c7_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 500 -p no:cacheprovider > gpurun_out/c7_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c7_pytest.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/c7_bench_1gpu.json 2> gpurun_out/c7_bench_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c7_bench_2gpu.json 2> gpurun_out/c7_bench_2gpu.err; echo "rc=$?" >> gpurun_out/c7_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --dp nccl --no-e2e --no-configs > gpurun_out/c7_bench_2gpu_nccl.json 2> gpurun_out/c7_bench_2gpu_nccl.err; echo "rc=$?" >> gpurun_out/c7_bench_2gpu_nccl.err
tail -4 gpurun_out/c7_pytest.txt
for f in c7_bench_1gpu c7_bench_2gpu c7_bench_2gpu_nccl; do python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), d['clocks'].get('samples'))
except Exception as e: print('$f', 'ERR', e)
"; done; tail -3 gpurun_out/c7_bench_2gpu.err
