#!/bin/bash
# block-manager mirror on a 16-CPU box: how many client threads, spin vs sleep, linger
cd "$(dirname "$0")/.."
O=gpurun_out
{
export GARAGE_BM_TRACE=1
run() { echo "== $*"; env $1 timeout 300 python tools/bm_bench.py ${@:2} 2>&1 | grep "garage_bm\|^{"; }
run X=1 --threads 16 --blocks 512 --no-verify
run X=1 --threads 24 --blocks 341 --no-verify
run X=1 --threads 32 --blocks 256 --no-verify
run GARAGE_BM_SPIN_WAIT=1 --threads 32 --blocks 256 --no-verify
run X=1 --threads 40 --blocks 205 --no-verify
run X=1 --threads 32 --blocks 256 --no-verify --linger-us 100
run X=1 --threads 32 --blocks 256
} 2>&1 | tee $O/r02_r18_bm.log
