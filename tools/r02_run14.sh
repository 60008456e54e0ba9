#!/bin/bash
# why is one dispatcher call 3-4x slower inside the loaded block manager than alone?
cd "$(dirname "$0")/.."
O=gpurun_out
{
echo "== default"; timeout 200 python tools/conc_probe.py --iters 150 2>&1 | tail -13
echo "== GARAGE_EC_BATCHCOPY=0"; GARAGE_EC_BATCHCOPY=0 timeout 200 python tools/conc_probe.py --iters 150 2>&1 | tail -13
echo "== 12 busy client threads"; timeout 200 python tools/conc_probe.py --iters 150 --busy 12 --threads 3 2>&1 | tail -5
echo "== trace, 3 threads, batch 16"; GARAGE_EC_TRACE=1 timeout 200 python tools/conc_probe.py --iters 150 --threads 3 --batch 16 2>&1 | tail -9
echo "== trace, 1 thread, batch 16"; GARAGE_EC_TRACE=1 timeout 200 python tools/conc_probe.py --iters 150 --threads 1 --batch 16 2>&1 | tail -5
} 2>&1 | tee $O/r02_r14_conc_probe.log
