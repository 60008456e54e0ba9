#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02_r10_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r10_pytest.log; tail -3 $O/r02_r10_pytest.log
timeout 300 python tools/sweep_bench.py --stripes 4096 > $O/r02_sweep_config5_adler8.json 2>&1; cat $O/r02_sweep_config5_adler8.json
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), {k: round(v["frac"], 3) for k, v in d["roofline"]["kernels"].items()}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
print("sweep", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in (d.get("config5_sweep") or {}).items() if isinstance(v, dict)})
print("e2e limiter:", d["e2e"]["limiter"])
PY
