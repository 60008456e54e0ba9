#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; V=build/variants
timeout 900 python -m pytest tests/test_block_manager.py tests/test_gpu_parity.py -x -q -m gpu > $O/r02_r4_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r4_pytest.log
S=$O/r02_r4_sweep.log; : > $S
for km in "10 4" "6 3" "4 2" "8 3" "12 4" "14 4" "16 4" "20 4" "24 4" "32 8" "7 3" "3 2"; do set -- $km
  timeout 120 python tools/kbench.py --k $1 --m $2 --tag default >> $S 2>&1
done
for v in d0x0 d1x0 d0x1; do for km in "10 4" "6 3"; do set -- $km
  timeout 120 python tools/kbench.py --so $V/libgarage_ec_$v.so --k $1 --m $2 --tag $v >> $S 2>&1
done; done
timeout 120 python tools/kbench.py --k 10 --m 4 --erasures 1 --same-pattern --tag default_same1 >> $S 2>&1
timeout 120 python tools/kbench.py --k 10 --m 4 --erasures 1 --tag default_e1 >> $S 2>&1
grep -h '^{' $S | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-14s k=%2d m=%d e=%d ok=%d enc %.3f dec %.3f ver %.3f' % (d['tag'], d['k'], d['m'], d['erasures'], d['ok'], d['encode_frac'], d['decode_frac'], d['verify_frac']))
"
for t in 16 64 128; do timeout 300 python tools/bm_bench.py --threads $t --blocks $((2048/t)) >> $O/r02_r4_bm.log 2>&1; done
timeout 300 python tools/bm_bench.py --threads 64 --blocks 32 --sum-kind 0 >> $O/r02_r4_bm.log 2>&1
timeout 300 python tools/bm_bench.py --threads 64 --blocks 32 --no-verify >> $O/r02_r4_bm.log 2>&1
cat $O/r02_r4_bm.log
timeout 600 python bench.py --steps 5 --blocks 1024 --sweep-stripes 512 --sweep-e2e-stripes 128 --cpu-blocks 64 > $O/r02_r4_bench.json 2> $O/r02_r4_bench.err; echo "bench rc=$?"
tail -3 $O/r02_r4_pytest.log; tail -5 $O/r02_r4_bench.err; cut -c1-3000 $O/r02_r4_bench.json
