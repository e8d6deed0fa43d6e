#!/bin/bash
# round 2, GPU call 23 (1 GPU): env write-back behind the phase-1 barrier; full GPU suite, smoke, bench, env trace
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c23_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c23_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c23_smoke.txt 2>&1
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c23_bench_1gpu.json 2> gpurun_out/c23_bench_1gpu.err
timeout 300 python bench.py --gpus 1 --envs 16384 $B > gpurun_out/c23_bench_16k.json 2> gpurun_out/c23_bench_16k.err
timeout 300 python bench.py --gpus 1 --envs 65536 --algo ddqn $B > gpurun_out/c23_bench_64k.json 2> gpurun_out/c23_bench_64k.err
UAVRL_ENV_TRACE=1 timeout 200 python bench.py --gpus 1 $B --min-seconds 0.05 2>&1 | grep "env_trace" | tail -3 > gpurun_out/c23_trace.txt
tail -4 gpurun_out/c23_pytest.txt; tail -2 gpurun_out/c23_smoke.txt
for f in c23_bench_1gpu c23_bench_16k c23_bench_64k; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; done
cat gpurun_out/c23_trace.txt
