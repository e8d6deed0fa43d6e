#!/bin/bash
# round 2, GPU call 19 (1 GPU): SAC persistent grid + parallel statistics (tests, loop timing); the driver's full 1-GPU bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sac_gpu.py tests/test_plugins_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c19_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c19_pytest.txt
timeout 300 python tools/sac_probe.py 16384 20 1048576 > gpurun_out/c19_sac.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c19_bench_1gpu.json 2> gpurun_out/c19_bench_1gpu.err
tail -4 gpurun_out/c19_pytest.txt; grep sac_probe gpurun_out/c19_sac.txt
python -c "
import json
d=json.loads(open('gpurun_out/c19_bench_1gpu.json').read().strip().splitlines()[-1]); print(round(d['value']/1e6,2),'M steps/s', round(d['ms_per_step']*1e3,2),'us/iter', 'e2e', round(d['e2e']['value']/1e6,2), {k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()}); print(d['roofline']); print({k:round(v['value']/1e6,1) for k,v in d['configs'].items()}); print(d['cpu_baseline'])"
