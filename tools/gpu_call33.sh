#!/bin/bash
# round 2, GPU call 33 (1 GPU): TD passes inside the training kernel on 64-row tiles (4 737 - 9 472 samples): tests + configs[3]'s per-GPU shape
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_tc_gpu.py tests/test_learner_gpu.py tests/test_multigpu_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c33_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c33_pytest.txt
B="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-configs"
timeout 300 python bench.py --gpus 1 --envs 8192 --algo ddqn $B > gpurun_out/c33_bench_8k.json 2> gpurun_out/c33_bench_8k.err
timeout 300 python bench.py --gpus 1 $B > gpurun_out/c33_bench_1gpu.json 2> gpurun_out/c33_bench_1gpu.err
tail -4 gpurun_out/c33_pytest.txt
for f in c33_bench_8k c33_bench_1gpu; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})"; done
