#!/bin/bash
# Round-end measurement sequence on one B200 (run through gpurun): everything lands in gpurun_out/r01_*.
# Numbers printed by runs under ncu are never used as bench values.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r01_pytest_gpu.txt
timeout 120 python -m pytest tests/test_env_gpu.py -m gpu -q -s -k device_pool 2>&1 | grep "pool of" >> $O/r01_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > $O/r01_smoke.txt 2>&1
timeout 400 python bench.py > $O/r01_bench_1gpu.json 2> $O/r01_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 200 --warmup 3 > $O/r01_bench_reference_arm.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/r01_launches.csv \
    python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --replay 65536 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_|env_kernel|reduce_adam" -s 60 -c 12 -f -o $O/r01_full_final \
    python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --replay 65536 > /dev/null 2>&1
UAVRL_TC_TRACE=1 timeout 100 python tools/tc_trace.py 2>&1 | grep "_trace" | tail -4 > $O/r01_trace.txt
b() { timeout 200 python bench.py --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(' '.join(sys.argv[1:]), '|', round(d['value']/1e6,2),'M steps/s |',round(d['ms_per_step']*1e3,2),'us/iter |',{k:round(v['ms']*1e3,1) for k,v in d['kernels'].items()})" "$@"; }
{
  b --envs 16384 --steps 600; b --envs 16384 --steps 600 --tc 0
  b --envs 65536 --steps 300; b --envs 65536 --steps 300 --tc 0
  b --steps 2000 --tc 0; b --steps 2000 --pdl 0; b --steps 2000 --per 1
  b --steps 2000 --algo ddqn; b --steps 2000 --algo dueling --net vanet2
} > $O/r01_variants.txt 2>&1
{ timeout 200 python tools/bench_sac.py; timeout 200 python tools/bench_sac.py --envs 4096; } 2>/dev/null | tail -2 > $O/r01_sac.txt
echo done
