#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_blake2.py tests/test_scrub_repair.py tests/test_block_manager.py tests/test_host_contract.py -x -q -m gpu > $O/r02_r9_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r9_pytest.log; tail -3 $O/r02_r9_pytest.log
timeout 300 python tools/sweep_bench.py --stripes 4096 > $O/r02_sweep_config5_adler8.json 2>&1; cat $O/r02_sweep_config5_adler8.json
timeout 300 python tools/sweep_bench.py --stripes 4096 --sum-kind 0 > $O/r02_sweep_config5_blake2.json 2>&1; cat $O/r02_sweep_config5_blake2.json
timeout 600 python tools/bm_bench.py --threads 64 --blocks 128 > $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 128 --blocks 64 >> $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 128 --blocks 64 --no-verify >> $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 16 --blocks 256 >> $O/r02_bm.log 2>&1
cat $O/r02_bm.log
timeout 900 python bench.py --no-cpu > $O/r02_r9_bench.json 2> $O/r02_r9_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02_r9_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), {k: round(v["frac"], 3) for k, v in d["roofline"]["kernels"].items()})
print("sweep", json.dumps(d.get("config5_sweep"))[:1800])
PY
