#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for z in 0 1; do
  GARAGE_EC_ZEROCOPY=$z timeout 600 python bench.py --no-cpu --no-sweep --steps 5 > $O/r02_r12_bench_z$z.json 2> $O/r02_r12_bench_z$z.err; echo "z=$z rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02_r12_bench_z$z.json").read().strip().splitlines() if l.startswith("{")][-1])
e = d["e2e"]; r = e["per_rank"][0]
print("zerocopy=$z e2e %.1f GiB/s checked=%s enc h2d %.1f d2h %.1f  rec h2d %.1f d2h %.1f" % (e["value"], e["checked"], r["encode_h2d_GBs"], r["encode_d2h_GBs"], r["reconstruct_h2d_GBs"], r["reconstruct_d2h_GBs"]))
PY
  tail -2 $O/r02_r12_bench_z$z.err
done
