#!/bin/bash
# split kernel over the whole GPU; is the GPU at low clocks while the block manager feeds it small batches?
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "block or Block or split or host" > $O/r02_r15_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_r15_pytest.log
{
echo "== conc probe"; GARAGE_EC_TRACE=1 timeout 200 python tools/conc_probe.py --iters 150 --threads 1 3 --batch 16 2>&1 | grep -v "^{\"busy" | tail -12
nvidia-smi --query-gpu=clocks.sm,clocks.mem,utilization.gpu,power.draw --format=csv,noheader,nounits -lms 50 > $O/r02_r15_clocks.csv &
SMI=$!
export GARAGE_BM_TRACE=1 GARAGE_EC_TRACE=1
for t in "64 128" "128 64"; do set -- $t
  echo "== bm_bench threads $1"; date +%s.%N
  timeout 300 python tools/bm_bench.py --threads $1 --blocks $2 --no-verify 2>&1 | grep -v "^garage_ec trace.*calls=0" | tail -16
  date +%s.%N
done
kill $SMI
python - <<PY
import statistics
rows=[tuple(float(x) for x in l.split(",")) for l in open("gpurun_out/r02_r15_clocks.csv") if l.strip()]
sm=[r[0] for r in rows]; print("clock samples", len(rows), "sm MHz min/median/max", min(sm), statistics.median(sm), max(sm), "mem", min(r[1] for r in rows), max(r[1] for r in rows))
print("distinct sm clocks:", sorted(set(sm)))
PY
} 2>&1 | tee $O/r02_r15_bm.log
