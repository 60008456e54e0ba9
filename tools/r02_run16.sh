#!/bin/bash
# landing / gather copies with non-temporal stores (the DMA read then comes from DRAM, not from a core's cache): A/B
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "block_manager or Block" > $O/r02_r16_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_r16_pytest.log
{
export GARAGE_BM_TRACE=1 GARAGE_EC_TRACE=1
for nt in 0 1; do for t in "64 128" "128 64"; do set -- $t
  echo "== GARAGE_BM_NT_COPY=$nt threads $1"
  GARAGE_BM_NT_COPY=$nt timeout 300 python tools/bm_bench.py --threads $1 --blocks $2 --no-verify 2>&1 | grep -v "calls=0" | grep -A1 "garage_bm\|^{\|ctx" | grep -v "^--" | head -9
done; done
} 2>&1 | tee $O/r02_r16_bm.log
