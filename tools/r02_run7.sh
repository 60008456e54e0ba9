#!/bin/bash
# final 1-GPU artefacts of round 2: bench lines, ncu metric summaries (raw pages as CSV: the .ncu-rep files are
# too large to bring back), launch list, sanitizer logs, block-manager load generator, call latency
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_reference_n1.json 2> $O/r02_bench_reference_n1.err
NCU="ncu --set full --clock-control none --kernel-name-base demangled --launch-skip 3 --launch-count 1 -f"
i=0
for name in encode reconstruct verify; do
  timeout 300 $NCU -k "regex:rs_apply_kernel<\(int\)10, \(int\)$i>" -o /tmp/ncu_$name python tools/kbench.py --k 10 --m 4 --blocks 4096 --iters 3 > $O/r02_ncu_$name.log 2>&1
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/r02_ncu_${name}_raw.csv 2>/dev/null
  i=$((i+1))
done
timeout 300 $NCU -k "regex:adler8_shards_kernel" -o /tmp/ncu_adler8 python tools/sweep_bench.py > $O/r02_ncu_adler8.log 2>&1
ncu -i /tmp/ncu_adler8.ncu-rep --page raw --csv > $O/r02_ncu_adler8_raw.csv 2>/dev/null
timeout 300 python tools/sweep_bench.py > $O/r02_sweep_config5_adler8.json 2>&1; timeout 300 python tools/sweep_bench.py --sum-kind 0 > $O/r02_sweep_config5_blake2.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:garage_ec -c 60 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-sweep > $O/r02_launches_bench.log 2>&1
( timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_small.py; timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_small.py; timeout 200 compute-sanitizer --tool synccheck python tools/sanitize_small.py ) > $O/r02_compute_sanitizer.log 2>&1
for km in "10 4" "6 3" "4 2" "17 4" "21 4" "25 4" "28 4" "32 4"; do set -- $km; timeout 120 python tools/kbench.py --k $1 --m $2 --tag final >> $O/r02_kbench_final_extra.log 2>&1; done
grep -h '^{' $O/r02_kbench_final_extra.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-8s k=%2d m=%d ok=%d enc %.3f dec %.3f ver %.3f' % (d['tag'], d['k'], d['m'], d['ok'], d['encode_frac'], d['decode_frac'], d['verify_frac']))
"
timeout 300 python tools/blocklat.py > $O/r02_blocklat.json 2>&1
for t in 16 64 128; do timeout 300 python tools/bm_bench.py --threads $t --blocks $((2048/t)) >> $O/r02_bm.log 2>&1; done
timeout 300 python tools/bm_bench.py --threads 128 --blocks 16 --no-verify >> $O/r02_bm.log 2>&1
cat $O/r02_sweep_config5_adler8.json $O/r02_sweep_config5_blake2.json; du -sh $O; grep -c "ERROR SUMMARY: 0 errors\|RACECHECK SUMMARY: 0 hazards" $O/r02_compute_sanitizer.log; tail -2 $O/r02_bench_n1.err; cat $O/r02_bm.log; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"))
print({k: (round(v["frac"], 3), round(v["avg_launch_ms"], 4)) for k, v in d["roofline"]["kernels"].items()}, d["roofline"]["kernel"], d["roofline"]["frac"])
print("cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
print("sweep", json.dumps(d.get("config5_sweep"))[:1500])
PY
