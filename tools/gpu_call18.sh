#!/bin/bash
# round 2, GPU call 18 (1 GPU): source-level ncu capture of the SAC tile kernels and of the DQN loop kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sac_target|sac_critic|sac_actor|sac_finish' -s 12 -c 4 -o gpurun_out/c18_sac_prof \
    python tools/sac_probe.py 16384 4 65536 > gpurun_out/c18_sac_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_|env_kernel|reduce_adam' -s 536 -c 10 -o gpurun_out/c18_prof \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --min-seconds 0.001 > gpurun_out/c18_ncu_full.log 2>&1
ls -la gpurun_out/c18_*; tail -2 gpurun_out/c18_sac_ncu.log
