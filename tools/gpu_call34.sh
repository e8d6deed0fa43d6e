#!/bin/bash
# round 2, GPU call 34 (1 GPU): the whole GPU suite + smoke on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/c34_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c34_pytest.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c34_smoke.txt 2>&1
tail -3 gpurun_out/c34_pytest.txt; tail -2 gpurun_out/c34_smoke.txt
