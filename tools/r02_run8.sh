#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02_r8_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r8_pytest.log; tail -4 $O/r02_r8_pytest.log
GARAGE_EC_TRACE=1 timeout 300 python tools/bm_bench.py --threads 16 --blocks 128 > $O/r02_r8_bm_trace.log 2>&1
GARAGE_EC_TRACE=1 timeout 300 python tools/bm_bench.py --threads 64 --blocks 32 >> $O/r02_r8_bm_trace.log 2>&1
timeout 300 python tools/bm_bench.py --threads 128 --blocks 16 --no-verify >> $O/r02_r8_bm_trace.log 2>&1
cat $O/r02_r8_bm_trace.log
timeout 300 python tools/e2e_probe.py --blocks 2048 > $O/r02_r8_e2e_probe.json 2>&1; cat $O/r02_r8_e2e_probe.json
timeout 300 python tools/sweep_bench.py --stripes 4096 > $O/r02_r8_sweep_adler8.json 2>&1; cat $O/r02_r8_sweep_adler8.json
