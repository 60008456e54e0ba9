#!/bin/bash
# 8-GPU validation: weak scaling of the device-resident step and of the HOST-buffer (e2e) step with NUMA-local pinned
# buffers, config 4 at 65 536 blocks, config-5 sweeps; which ranks slow each other down (e2e_probe); the block-manager
# load generator on the same host
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi topo -m > $O/r02_n8_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_n8_bench.json 2> $O/r02_n8_bench.err; echo "n8 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29517 tools/e2e_probe.py --blocks 2048 > $O/r02_n8_e2e_probe.json 2> $O/r02_n8_e2e_probe.err; echo "probe rc=$?"
timeout 300 python tools/bm_bench.py --threads 64 --blocks 128 > $O/r02_n8host_bm.log 2>&1
timeout 300 python tools/bm_bench.py --threads 128 --blocks 64 >> $O/r02_n8host_bm.log 2>&1
timeout 300 python tools/bm_bench.py --threads 128 --blocks 64 --no-verify >> $O/r02_n8host_bm.log 2>&1
python - <<'PY'
import json
for f in ("r02_n8_bench.json",):
    try:
        d = json.loads([l for l in open("gpurun_out/" + f).read().strip().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    e = d.get("e2e") or {}
    print(f, "value", round(d["value"]), "e2e", e.get("value"), "workload", d["config"]["workload"][:90])
    for r in e.get("per_rank", []):
        print("   rank", r["rank"], "GiB/s %.1f" % r["GiBs"], "enc h2d %.1f" % r["encode_h2d_GBs"], "rec h2d %.1f" % r["reconstruct_h2d_GBs"], "gpu node", r["gpu_numa_node"], "buf node", r["pinned_buffer_numa_node"], "cpus", r["thread_affinity_cpus"])
    print("   roofline", {k: round(v["frac"], 3) for k, v in d["roofline"]["kernels"].items()}, "sweep", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in (d.get("config5_sweep") or {}).items() if isinstance(v, dict)})
PY
cat $O/r02_n8_e2e_probe.json | grep -v "^NCCL" | head -80; cat $O/r02_n8host_bm.log; tail -3 $O/r02_n8_bench.err
