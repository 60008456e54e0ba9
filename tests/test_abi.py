"""CPU-only checks of the drop-in boundary: libgarage_ec.so loads, exports every symbol
include/garage_ec.h declares, and -- with no GPU in the container -- refuses to create a
context instead of falling back to CPU code."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import garage_b200 as G  # noqa: E402


def _has_cuda():
    import torch

    return torch.cuda.is_available()


def header_decls():
    src = open(os.path.join(ROOT, "include", "garage_ec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(garage_ec_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_decls() == sorted(G.ABI_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    L = G.load_library()
    for name in header_decls():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", G.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (garage_ec_[a-z0-9_]+)", out))
    assert set(header_decls()) <= exported
    assert L.garage_ec_abi_version() == 1


def test_built_for_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", G.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_geometry_helpers_no_gpu_needed():
    L = G.load_library()
    assert L.garage_ec_shard_len(1 << 20, 10) == 104858
    assert L.garage_ec_shard_len(1 << 20, 6) == 174763
    assert L.garage_ec_shard_len(1 << 20, 4) == 262144
    assert L.garage_ec_shard_len(0, 4) == 0
    assert L.garage_ec_shard_len(1, 4) == 1
    assert L.garage_ec_stride_for(104858) == 104960
    assert L.garage_ec_stride_for(262144) == 262144
    assert L.garage_ec_stride_for(1) == 128


def test_strerror():
    L = G.load_library()
    for code in range(0, -7, -1):
        assert L.garage_ec_strerror(code)
    assert b"fallback" in L.garage_ec_strerror(G.E_NODEVICE)


def test_bad_arguments_rejected_before_touching_cuda():
    L = G.load_library()
    h = C.c_void_p()
    assert L.garage_ec_create(C.byref(h), 0, 0, 4, 0) == G.E_INVALID
    assert L.garage_ec_create(C.byref(h), 0, 33, 4, 0) == G.E_INVALID
    assert L.garage_ec_create(C.byref(h), 0, 10, 9, 0) == G.E_INVALID
    assert L.garage_ec_create(C.byref(h), 0, 10, 4, 7) == G.E_INVALID
    assert L.garage_ec_create_with_matrix(C.byref(h), 0, 10, 4, None) == G.E_INVALID
    assert not h.value
    assert L.garage_ec_encode(None, None, None, None, 16, 1, 0, None) == G.E_INVALID


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU behaviour")
def test_no_device_means_error_not_cpu_fallback():
    with pytest.raises(G.EcError) as ei:
        G.GarageEc(device=0, k=10, m=4)
    assert ei.value.code == G.E_NODEVICE


def test_product_does_not_touch_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "garage_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                for pat in (r'#\s*include\s*["<][^">]*oracle', r"^\s*(import|from)\s+\S*oracle",
                            r"librs_oracle", r"oracle_lib", r"rs_(oracle|simd)_\w+\s*\("):
                    assert not re.search(pat, txt, flags=re.M), (f, pat)
    out = subprocess.run(["ldd", G.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_rust_shim_declares_every_symbol():
    """integration/garage_block_cuda/src/sys.rs (source only: no rustc here) must bind exactly
    the symbols the header declares"""
    src = open(os.path.join(ROOT, "integration", "garage_block_cuda", "src", "sys.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (garage_ec_[a-z0-9_]+)\s*\(", src)))
    assert rust == header_decls()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert sorted(set(re.findall(r"pub fn (garage_ec_[a-z0-9_]+)\s*\(", md))) == header_decls()


def test_copy_for_dma_is_a_memcpy_for_every_size_and_alignment():
    """garage_ec_copy_for_dma (non-temporal landing copy) moves exactly the bytes memcpy would: heads, tails,
    unaligned sources and destinations, sizes around the 4 KiB switch-over and the 64-byte body granule"""
    import numpy as np

    import garage_b200 as G

    rng = np.random.default_rng(12)
    src_buf = rng.integers(0, 256, (1 << 21) + 256, dtype=np.uint8)
    for n in (0, 1, 63, 64, 4095, 4096, 4097, 4096 + 63, 8191, 104858, (1 << 20), (1 << 20) + 17):
        for so in (0, 1, 15, 33):
            for do in (0, 1, 16, 63):
                dst_buf = np.full(n + 256, 0xA5, dtype=np.uint8)
                G.copy_for_dma(dst_buf[do:do + n], src_buf[so:so + n])
                assert np.array_equal(dst_buf[do:do + n], src_buf[so:so + n]), (n, so, do)
                assert (dst_buf[:do] == 0xA5).all() and (dst_buf[do + n:] == 0xA5).all(), (n, so, do)
