#!/bin/bash
# round 2, GPU call 4 (1 GPU): full tests (APF, persistent dW), bench, large-batch variants
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/c4_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c4_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?" >> gpurun_out/c4_bench.err
rm -f gpurun_out/c4_variants.txt
for v in "--envs 16384" "--envs 65536 --algo ddqn" "--envs 16384 --net vanet2 --algo dueling"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.3 $v 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(sys.argv[1:], round(d['value']/1e6,2),'M steps/s',round(d['ms_per_step']*1e3,2),'us/iter',{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" $v >> gpurun_out/c4_variants.txt 2>&1
done
tail -8 gpurun_out/c4_pytest.txt; head -c 300 gpurun_out/c4_bench.json; tail -2 gpurun_out/c4_bench.err; cat gpurun_out/c4_variants.txt
