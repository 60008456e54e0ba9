#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into one JSON: the metrics DESIGN.md / profiles cite.

    python tools/ncu_summary.py out.json name1=file1.ncu-rep name2=file2.ncu-rep ...
"""
import csv
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
]


def one(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"Kernel Name": vals[hdr.index("Kernel Name")]}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            d[w] = ("%s %s" % (vals[i], units[i])).strip()
    return d


def _bytes(txt):
    """'4.299205 Gbyte' -> bytes"""
    v, u = txt.split()[:2]
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
    return float(v) * mul


def traffic_json(summary_path, out_path, blocks, note):
    """profiles/dominant_kernel_traffic.json (read by bench.py): dram read+write bytes per launch of the three
    streaming kernels from captures named encode / reconstruct / verify in a summary written by this tool"""
    d = json.load(open(summary_path))
    res = {"note": note, "source": summary_path, "kernels": {}}
    for name in ("encode", "reconstruct", "verify"):
        if name in d:
            r, w = _bytes(d[name]["dram__bytes_read.sum"]), _bytes(d[name]["dram__bytes_write.sum"])
            res["kernels"][name] = {"dram_bytes_per_launch": r + w, "dram_bytes_read": r, "dram_bytes_write": w,
                                    "blocks": blocks, "kernel": d[name]["Kernel Name"],
                                    "gpu_time_under_ncu": d[name].get("gpu__time_duration.sum")}
    json.dump(res, open(out_path, "w"), indent=1)
    return res


if __name__ == "__main__":
    res = {}
    for spec in sys.argv[2:]:
        name, path = spec.split("=", 1)
        res[name] = one(path)
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    for k, v in res.items():
        print(k, v.get("gpu__time_duration.sum"), v.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
