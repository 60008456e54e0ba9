#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; V=build/variants
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_scrub_repair.py -x -q -m gpu > $O/r02_r6_pytest.log 2>&1; echo "rc=$?" >> $O/r02_r6_pytest.log
S=$O/r02_r6_sweep.log; : > $S
kb() { timeout 120 python tools/kbench.py "$@" >> $S 2>&1; }
for km in "10 4" "6 3" "4 2" "8 3" "12 4" "7 3" "3 2" "5 2" "9 3" "11 4" "13 4" "14 4" "16 4" "17 4" "20 4" "24 4" "28 4" "32 4" "32 8" "17 8" "2 8"; do set -- $km; kb --k $1 --m $2 --tag default; done
kb --k 10 --m 4 --erasures 1 --tag default_e1; kb --k 10 --m 4 --erasures 2 --tag default_e2; kb --k 10 --m 4 --erasures 1 --same-pattern --tag default_same1
grep -h '^{' $S | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-16s k=%2d m=%d e=%d ok=%d enc %.3f dec %.3f ver %.3f' % (d['tag'], d['k'], d['m'], d['erasures'], d['ok'], d['encode_frac'], d['decode_frac'], d['verify_frac']))
"
timeout 900 python bench.py > $O/r02_r6_bench.json 2> $O/r02_r6_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_r6_bench_reference.json 2> $O/r02_r6_bench_reference.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled --launch-skip 3 --launch-count 1 -f"
for mode in 0 1 2; do
  timeout 300 $NCU -k "regex:rs_apply_kernel<\(int\)10, \(int\)$mode>" -o $O/r02_ncu_final_mode$mode python tools/kbench.py --k 10 --m 4 --blocks 4096 --iters 3 > $O/r02_ncu_final_mode$mode.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:garage_ec -c 60 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-sweep > $O/r02_launches_bench.log 2>&1
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_small.py > $O/r02_r6_memcheck.log 2>&1
timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_small.py > $O/r02_r6_racecheck.log 2>&1
timeout 200 compute-sanitizer --tool synccheck python tools/sanitize_small.py > $O/r02_r6_synccheck.log 2>&1
tail -3 $O/r02_r6_pytest.log; tail -2 $O/r02_r6_memcheck.log $O/r02_r6_racecheck.log $O/r02_r6_synccheck.log; tail -3 $O/r02_r6_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_r6_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"))
print({k: (round(v["frac"], 3), round(v["avg_launch_ms"], 4)) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["kernel"], d["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("sweep", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in (d.get("config5_sweep") or {}).items() if isinstance(v, dict)})
print("clocks", d["clocks"])
PY
