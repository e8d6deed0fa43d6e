#!/bin/bash
# round 2, GPU call 32 (1 GPU): final code at the larger configurations, for the DESIGN table
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 --envs 16384 $B > gpurun_out/c32_bench_16k.json 2> gpurun_out/c32_bench_16k.err
timeout 300 python bench.py --gpus 1 --envs 65536 --algo ddqn $B > gpurun_out/c32_bench_64k.json 2> gpurun_out/c32_bench_64k.err
timeout 300 python bench.py --gpus 1 --envs 8192 --algo ddqn $B > gpurun_out/c32_bench_8k.json 2> gpurun_out/c32_bench_8k.err
timeout 300 python tools/sac_probe.py 16384 20 1048576 > gpurun_out/c32_sac.txt 2>&1
for f in c32_bench_16k c32_bench_64k c32_bench_8k; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}, d['roofline']['kernel'], round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))"; done; grep sac_probe gpurun_out/c32_sac.txt
