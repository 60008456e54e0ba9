#!/bin/bash
# validation of the round's final code: full GPU suite, default bench, dispatcher-call probe, block-manager mirror
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_r17_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_r17_pytest.log
timeout 600 python bench.py > $O/r02_r17_bench.json 2> $O/r02_r17_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_r17_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
e = d["e2e"]
print("value %.1f e2e %.1f GiB/s checked=%s cpu %.1f launches %s" % (d["value"], e["value"], e["checked"], d["cpu_baseline"]["value"], d["gpu_launches"]))
PY
{
echo "== conc probe"; timeout 200 python tools/conc_probe.py --iters 150 --threads 1 3 --batch 16 64 2>&1 | grep "^{'" 
export GARAGE_BM_TRACE=1
for cfg in "32 256 --no-verify" "48 171 --no-verify" "64 128 --no-verify" "96 86 --no-verify" "64 128"; do set -- $cfg
  echo "== bm_bench threads $1 $3"
  timeout 300 python tools/bm_bench.py --threads $1 --blocks $2 $3 2>&1 | grep "garage_bm\|^{"
done
echo "== bm_bench threads 64 --no-verify, dispatchers spin (GARAGE_BM_SPIN_WAIT=1)"
GARAGE_BM_SPIN_WAIT=1 timeout 300 python tools/bm_bench.py --threads 64 --blocks 128 --no-verify 2>&1 | grep "garage_bm\|^{"
} 2>&1 | tee $O/r02_r17_bm.log
