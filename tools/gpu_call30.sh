#!/bin/bash
# round 2, GPU call 30 (2 GPUs): final code -- 2-rank tests and the driver's exact commands at N = 1 and N = 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c30_bench_1gpu.json 2> gpurun_out/c30_bench_1gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c30_bench_2gpu.json 2> gpurun_out/c30_bench_2gpu.err; echo "rc=$?" >> gpurun_out/c30_bench_2gpu.err
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c30_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c30_pytest.txt
for f in c30_bench_1gpu c30_bench_2gpu; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d.get('e2e',{}).get('value',0)/1e6,2), 'blocks', d.get('repeats'), round(d.get('block_ms_min',0),2), round(d.get('block_ms_median',0),2), round(d.get('block_ms_max',0),2))"; done; tail -2 gpurun_out/c30_pytest.txt
