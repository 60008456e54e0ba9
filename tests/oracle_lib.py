"""ctypes loader for the CPU oracle (oracle/librs_oracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs -- never by garage_b200."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "librs_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)


def build(force=False):
    srcs = [os.path.join(ODIR, f) for f in ("rs_oracle.c", "rs_simd.c", "rs_oracle.h")]
    multi = int(os.environ.get("WORLD_SIZE", "1")) > 1  # only rank 0 uses the oracle; never race on it
    stale = os.path.exists(SO) and any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs)
    if force or not os.path.exists(SO) or (stale and not multi):
        if all(os.path.exists(s) for s in srcs):
            subprocess.run(["make", "-C", ODIR, "-s", "-B"], check=True)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.rs_oracle_gf_mul.restype = C.c_uint8
        L.rs_oracle_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
        L.rs_oracle_gf_inv.restype = C.c_uint8
        L.rs_oracle_gf_inv.argtypes = [C.c_uint8]
        L.rs_oracle_gf_exp.restype = C.c_uint8
        L.rs_oracle_gf_exp.argtypes = [C.c_int]
        L.rs_oracle_gf_log.restype = C.c_uint8
        L.rs_oracle_gf_log.argtypes = [C.c_uint8]
        L.rs_oracle_build_matrix.restype = C.c_int
        L.rs_oracle_build_matrix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.rs_oracle_invert.restype = C.c_int
        L.rs_oracle_invert.argtypes = [C.c_void_p, C.c_int]
        geo = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.rs_oracle_encode.restype = None
        L.rs_oracle_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + geo
        L.rs_oracle_reconstruct.restype = C.c_size_t
        L.rs_oracle_reconstruct.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p] + geo
        L.rs_oracle_verify.restype = None
        L.rs_oracle_verify.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + geo
        L.rs_oracle_shard_len.restype = C.c_uint32
        L.rs_oracle_shard_len.argtypes = [C.c_uint32, C.c_int]
        L.rs_oracle_split_block.restype = None
        L.rs_oracle_split_block.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_size_t]
        L.rs_oracle_join_block.restype = None
        L.rs_oracle_join_block.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_size_t, C.c_void_p]
        L.rs_oracle_fill_random.restype = None
        L.rs_oracle_fill_random.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64]
        L.rs_simd_encode.restype = None
        L.rs_simd_encode.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + geo + [C.c_int]
        L.rs_simd_reconstruct.restype = C.c_size_t
        L.rs_simd_reconstruct.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p] + geo + [C.c_int]
        L.rs_simd_verify.restype = None
        L.rs_simd_verify.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + geo + [C.c_int]
        L.rs_simd_parallel_copy.restype = None
        L.rs_simd_parallel_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
        L.rs_simd_isa.restype = C.c_char_p
        L.rs_simd_max_threads.restype = C.c_int
        L.rs_simd_force_isa.restype = C.c_int
        L.rs_simd_force_isa.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_matrix(k, m, kind=0):
    P = np.zeros((m, k), dtype=np.uint8)
    rc = lib().rs_oracle_build_matrix(k, m, kind, _p(P))
    if rc:
        raise ValueError("bad (k,m,kind)")
    return P


def _lens(shard_len, n):
    if shard_len is None:
        return None
    a = np.ascontiguousarray(shard_len, dtype=np.uint32)
    assert a.shape == (n,)
    return a


def encode(k, m, P, data, stride, n, shard_len=None, simd=False, threads=0):
    """data: flat uint8 n*k*stride -> parity flat n*m*stride (zero-initialised)."""
    parity = np.zeros(n * m * stride, dtype=np.uint8)
    sl = _lens(shard_len, n)
    P = np.ascontiguousarray(P, dtype=np.uint8)
    if simd:
        lib().rs_simd_encode(k, m, _p(P), _p(data), _p(parity), _p(sl), stride, n, threads)
    else:
        lib().rs_oracle_encode(k, m, _p(P), _p(data), _p(parity), _p(sl), stride, n)
    return parity


def reconstruct(k, m, P, shards, present, stride, n, shard_len=None, simd=False, threads=0):
    """shards modified in place; returns (n_bad, status)."""
    status = np.zeros(n, dtype=np.int32)
    sl = _lens(shard_len, n)
    P = np.ascontiguousarray(P, dtype=np.uint8)
    present = np.ascontiguousarray(present, dtype=np.uint8)
    if simd:
        bad = lib().rs_simd_reconstruct(k, m, _p(P), _p(shards), _p(present), _p(status), _p(sl),
                                        stride, n, threads)
    else:
        bad = lib().rs_oracle_reconstruct(k, m, _p(P), _p(shards), _p(present), _p(status),
                                          _p(sl), stride, n)
    return bad, status


def verify(k, m, P, shards, stride, n, shard_len=None, simd=False, threads=0):
    mm = np.zeros(n, dtype=np.uint32)
    sl = _lens(shard_len, n)
    P = np.ascontiguousarray(P, dtype=np.uint8)
    if simd:
        lib().rs_simd_verify(k, m, _p(P), _p(shards), _p(mm), _p(sl), stride, n, threads)
    else:
        lib().rs_oracle_verify(k, m, _p(P), _p(shards), _p(mm), _p(sl), stride, n)
    return mm


def numa_local_copy(a, bytes_per_stripe, n, threads=0):
    """copy of `a` whose pages are first touched by the worker threads that will process them"""
    out = np.empty_like(a)
    lib().rs_simd_parallel_copy(_p(out), _p(a), bytes_per_stripe, n, threads)
    return out


def fill_random(n, seed, offset=0):
    a = np.empty(n, dtype=np.uint8)
    lib().rs_oracle_fill_random(_p(a), n, seed, offset)
    return a


def split_block(block, k, stride):
    block = np.ascontiguousarray(block, dtype=np.uint8)
    out = np.zeros(k * stride, dtype=np.uint8)
    lib().rs_oracle_split_block(_p(block), len(block), k, _p(out), stride)
    return out


def join_block(shards, block_len, k, stride):
    out = np.zeros(block_len, dtype=np.uint8)
    lib().rs_oracle_join_block(_p(shards), block_len, k, stride, _p(out))
    return out
