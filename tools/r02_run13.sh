#!/bin/bash
# final validation of round 2 + block-manager PUT trace: full GPU suite, default bench, mirror bench
cd "$(dirname "$0")/.."
O=gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_r13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_r13_pytest.log
timeout 600 python bench.py > $O/r02_r13_bench.json 2> $O/r02_r13_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_r13_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
e = d["e2e"]; r = e["per_rank"][0]
print("value %.1f e2e %.1f GiB/s checked=%s enc h2d %.1f d2h %.1f  rec h2d %.1f d2h %.1f  cpu %s" % (d["value"], e["value"], e["checked"], r["encode_h2d_GBs"], r["encode_d2h_GBs"], r["reconstruct_h2d_GBs"], r["reconstruct_d2h_GBs"], d["cpu_baseline"]["value"]))
print(json.dumps(d["roofline"]["kernels"]))
PY
export GARAGE_BM_TRACE=1
for t in "64 128" "128 64" "32 256"; do set -- $t
  timeout 300 python tools/bm_bench.py --threads $1 --blocks $2 --no-verify 2>&1 | tail -3
done 2>&1 | tee $O/r02_r13_bm.log
