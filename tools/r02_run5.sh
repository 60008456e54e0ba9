#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; V=build/variants
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_contract.py tests/test_block_manager.py tests/test_external_anchors.py -x -q -m gpu > $O/r02_r5_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r5_pytest.log
S=$O/r02_r5_sweep.log; : > $S
kb() { timeout 120 python tools/kbench.py "$@" >> $S 2>&1; }
for km in "10 4" "6 3" "4 2" "8 3" "12 4" "7 3" "3 2" "5 2" "9 3" "14 4" "16 4" "20 4" "24 4" "32 8"; do set -- $km; kb --k $1 --m $2 --tag default; done
kb --so $V/libgarage_ec_tm0.so --k 10 --m 4 --tag tmap0; kb --so $V/libgarage_ec_tm0.so --k 6 --m 3 --tag tmap0
kb --so $V/libgarage_ec_b12_nw20.so --k 12 --m 4 --tag b12_nw20
for K2 in 14 16 20 24; do for nw in 12 16 20; do kb --so $V/libgarage_ec_b${K2}_nw${nw}.so --k $K2 --m 4 --tag b${K2}_nw${nw}; done; done
for nw in 12 16 20; do kb --so $V/libgarage_ec_b32_nw${nw}.so --k 32 --m 8 --tag b32_nw${nw}; done
for nw in 20 24 28; do kb --so $V/libgarage_ec_v4t_nw${nw}.so --k 4 --m 2 --tag v4t_nw${nw}; done
kb --k 4 --m 2 --stride-pad 128 --tag default_pad128; kb --k 4 --m 2 --stride-pad 2048 --tag default_pad2048
kb --k 10 --m 4 --erasures 1 --tag default_e1; kb --k 10 --m 4 --erasures 2 --tag default_e2
grep -h '^{' $S | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-16s k=%2d m=%d e=%d ok=%d enc %.3f dec %.3f ver %.3f' % (d['tag'], d['k'], d['m'], d['erasures'], d['ok'], d['encode_frac'], d['decode_frac'], d['verify_frac']))
"
timeout 300 python tools/blocklat.py > $O/r02_r5_blocklat.json 2>&1; cat $O/r02_r5_blocklat.json
timeout 300 python tools/bm_bench.py --threads 64 --blocks 32 > $O/r02_r5_bm.log 2>&1; timeout 300 python tools/bm_bench.py --threads 128 --blocks 16 --no-verify >> $O/r02_r5_bm.log 2>&1; cat $O/r02_r5_bm.log
tail -3 $O/r02_r5_pytest.log
