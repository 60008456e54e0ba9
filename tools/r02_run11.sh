#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_block_manager.py tests/test_host_contract.py -x -q -m gpu > $O/r02_r11_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r11_pytest.log; tail -3 $O/r02_r11_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
: > $O/r02_bm.log
timeout 600 python tools/bm_bench.py --threads 64 --blocks 128 >> $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 128 --blocks 64 >> $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 128 --blocks 64 --no-verify >> $O/r02_bm.log 2>&1
timeout 600 python tools/bm_bench.py --threads 32 --blocks 256 --no-verify >> $O/r02_bm.log 2>&1
cat $O/r02_bm.log
