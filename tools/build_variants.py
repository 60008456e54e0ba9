#!/usr/bin/env python
"""Build tuning variants of libgarage_ec.so into build/variants/ (git-ignored, travels with gpurun).

    python tools/build_variants.py name=-DGEC_THREADS=384,-DGEC_PIPELINE=1 ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garage_b200 import _build  # noqa: E402

out_dir = os.path.join(ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
from concurrent.futures import ThreadPoolExecutor  # noqa: E402


def run(spec):
    name, flags = spec.split("=", 1)
    out = os.path.join(out_dir, "libgarage_ec_%s.so" % name)
    cmd = _build.nvcc_cmd(out=out, extra=tuple(f for f in flags.split(",") if f) + ("-Xptxas", "-v"))
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return name, out, p


with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2) - 1)) as ex:
    results = list(ex.map(run, sys.argv[1:]))
for name, out, p in results:
    txt = p.stdout
    if p.returncode:
        print(name, "FAILED\n", txt)
        continue
    lines = txt.splitlines()
    for i, l in enumerate(lines):
        if "rs_apply_kernelILi10ELi0" in l and "Compiling" in l:
            print(name, "enc10:", " | ".join(x.strip() for x in lines[i + 1:i + 4] if "registers" in x or "spill" in x))
