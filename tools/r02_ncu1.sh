#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; V=build/variants
timeout 900 python -m pytest tests/test_host_contract.py tests/test_blake2.py tests/test_external_anchors.py tests/test_raid6_anchor.py tests/test_scrub_repair.py -x -q -m gpu > $O/r02_n1_pytest.log 2>&1; echo "rc=$?" >> $O/r02_n1_pytest.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled --launch-skip 3 --launch-count 1 -f"
cap() { # name so k m mode
  local so=""; [ "$2" != default ] && so="--so $V/libgarage_ec_$2.so"
  timeout 300 $NCU -k "regex:rs_apply_kernel<\(int\)$3, \(int\)$5>" -o $O/r02_ncu_$1 python tools/kbench.py $so --k $3 --m $4 --blocks 2048 --iters 3 > $O/r02_ncu_$1.log 2>&1
}
cap ver10_tma16 t6_nw16 10 4 2
cap dec10_tma20 t6_nw20 10 4 1
cap enc10_tma16 t6_nw16 10 4 0
cap ver6_ldg16 l6_nw16 6 3 2
cap ver6_tma20 t6_nw20 6 3 2
cap enc10_ldg16 l6_nw16 10 4 0
timeout 600 python bench.py --steps 5 --blocks 1024 --sweep-stripes 512 --sweep-e2e-stripes 128 --cpu-blocks 64 > $O/r02_n1_bench.json 2> $O/r02_n1_bench.err; echo "bench rc=$?"
tail -3 $O/r02_n1_pytest.log; ls -la $O/*.ncu-rep; tail -5 $O/r02_n1_bench.err; cut -c1-1500 $O/r02_n1_bench.json
