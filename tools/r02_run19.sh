#!/bin/bash
# final defaults (spin wait, linger 100 us): block-manager tests + the mirror bench at its defaults, twice
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "block_manager" > $O/r02_r19_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 $O/r02_r19_pytest.log
{
export GARAGE_BM_TRACE=1
for i in 1 2; do timeout 100 python tools/bm_bench.py --no-verify 2>&1 | grep "garage_bm\|^{"; done
timeout 100 python tools/bm_bench.py 2>&1 | grep "garage_bm\|^{"
} 2>&1 | tee $O/r02_r19_bm.log
