#!/bin/bash
# round-2 second GPU pass: (TMA | LDG) x warps sweep for k = 4,6,8,10,12 and ncu --set full of the k = 10 kernels
cd "$(dirname "$0")/.."
O=gpurun_out
S=$O/r02_s2_sweep.log; : > $S
V=build/variants
run() { timeout 120 python tools/kbench.py --so $V/libgarage_ec_$1.so --k $2 --m $3 --tag $1 >> $S 2>&1; }
for nw in 8 12 16 20 24 28 32; do
  run t6_nw$nw 10 4; run l6_nw$nw 10 4
  run t4_nw$nw 4 2; run l4_nw$nw 4 2
  run t6_nw$nw 6 3; run l6_nw$nw 6 3
  run t8_nw$nw 8 3; run l8_nw$nw 8 3
  run t12_nw$nw 12 4; run l12_nw$nw 12 4
done
run llg2_nw16 10 4; run llg2_nw16 6 3
timeout 60 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r02_s2_racecheck.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled --launch-skip 3 --launch-count 1 -f"
timeout 300 $NCU -k 'regex:rs_apply_kernel<10, 0>' -o $O/r02_ncu_enc_ldg python tools/kbench.py --k 10 --m 4 --blocks 2048 --iters 3 > $O/r02_ncu_1.log 2>&1
timeout 300 $NCU -k 'regex:rs_apply_kernel<10, 1>' -o $O/r02_ncu_dec_ldg python tools/kbench.py --k 10 --m 4 --blocks 2048 --iters 3 > $O/r02_ncu_2.log 2>&1
timeout 300 $NCU -k 'regex:rs_apply_kernel<10, 2>' -o $O/r02_ncu_ver_tma python tools/kbench.py --k 10 --m 4 --blocks 2048 --iters 3 > $O/r02_ncu_3.log 2>&1
timeout 300 $NCU -k 'regex:rs_apply_kernel<10, 0>' -o $O/r02_ncu_enc_tma python tools/kbench.py --so $V/libgarage_ec_t6_nw16.so --k 10 --m 4 --blocks 2048 --iters 3 > $O/r02_ncu_4.log 2>&1
timeout 300 $NCU -k 'regex:rs_apply_kernel<10, 1>' -o $O/r02_ncu_dec_tma python tools/kbench.py --so $V/libgarage_ec_t6_nw16.so --k 10 --m 4 --blocks 2048 --iters 3 > $O/r02_ncu_5.log 2>&1
grep -h '^{' $S | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-10s k=%2d ok=%d enc %.3f dec %.3f ver %.3f' % (d['tag'], d['k'], d['ok'], d['encode_frac'], d['decode_frac'], d['verify_frac']))
"
tail -3 $O/r02_s2_racecheck.log; ls -la $O/*.ncu-rep
