#!/bin/bash
# round-2 first GPU pass: parity tests of the new streaming kernels, sanitizer on a tiny case, variant sweep
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/r02_s1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r02_s1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02_s1_pytest.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r02_s1_memcheck.log 2>&1; echo "rc=$?" >> $O/r02_s1_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r02_s1_racecheck.log 2>&1; echo "rc=$?" >> $O/r02_s1_racecheck.log
S=$O/r02_s1_sweep.log; : > $S
for v in default alltma allldg tma20 tma12; do
  so=""; [ $v != default ] && so="--so build/variants/libgarage_ec_$v.so"
  for km in "10 4" "6 3"; do set -- $km
    timeout 120 python tools/kbench.py $so --k $1 --m $2 --tag $v >> $S 2>&1
  done
done
for km in "4 2" "8 3" "12 4" "14 4" "16 4" "20 4" "24 4" "32 8" "7 3" "3 2"; do set -- $km
  timeout 120 python tools/kbench.py --k $1 --m $2 --tag default >> $S 2>&1
done
timeout 120 python tools/kbench.py --k 10 --m 4 --erasures 1 --same-pattern --tag default_same1 >> $S 2>&1
timeout 120 python tools/kbench.py --k 10 --m 4 --erasures 1 --tag default_e1 >> $S 2>&1
grep -h '^{' $S | cut -c1-400
tail -3 $O/r02_s1_pytest.log
