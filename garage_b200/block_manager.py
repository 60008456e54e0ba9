"""ctypes view of libgarage_block.so -- the C++ host-side mirror of Garage's BlockManager for
the erasure-coded path (include/garage_block_manager.h).  Harness only; the logic is C++."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _build

OK = 0
E_CORRUPT_DATA, E_MISSING_BLOCK, E_MESSAGE, E_QUORUM = -10, -11, -12, -13

SYMBOLS = [
    "garage_bm_default_config", "garage_bm_create", "garage_bm_destroy", "garage_bm_blake2sum",
    "garage_bm_rpc_put_block", "garage_bm_rpc_get_block", "garage_bm_resync_block", "garage_bm_resync_all",
    "garage_bm_repair_enqueue_missing", "garage_bm_scrub", "garage_bm_set_node_up", "garage_bm_corrupt_shard",
    "garage_bm_drop_shard", "garage_bm_node_shard_index", "garage_bm_storage_nodes_of", "garage_bm_get_metrics",
    "garage_bm_block_incref", "garage_bm_block_decref", "garage_bm_get_block_rc", "garage_bm_scrub_step",
    "garage_bm_bench", "garage_bm_corrupt_shard_header", "garage_bm_plant_stale_shard", "garage_bm_set_node_readonly",
]
SUM_BLAKE2, SUM_ADLER8 = 0, 1


class Config(C.Structure):
    _fields_ = [("data_shards", C.c_int), ("parity_shards", C.c_int), ("cuda_device", C.c_int),
                ("n_nodes", C.c_int), ("block_size", C.c_uint32), ("block_ram_buffer_max", C.c_uint64),
                ("batch_max_blocks", C.c_uint32), ("batch_linger_us", C.c_uint32), ("data_dir", C.c_char_p),
                ("shard_sum_kind", C.c_int), ("verify_content_hash", C.c_int), ("data_fsync", C.c_int),
                ("block_gc_delay_ms", C.c_uint32)]


class Metrics(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes_written", "bytes_read", "corruption_counter", "resync_counter", "resync_error_counter",
        "resync_recv_counter", "delete_counter", "put_calls", "put_batches", "reconstruct_calls",
        "reconstruct_batches", "scrub_shards_checked", "scrub_corruptions", "resync_queue_length",
        "encode_call_us", "reconstruct_call_us", "corrupt_data_errors", "write_errors")]


_lib = None


def load_library():
    global _lib
    if _lib is None:
        _build.build_block_manager()
        L = C.CDLL(_build.BM_SO)
        vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
        L.garage_bm_default_config.argtypes = [C.POINTER(Config)]
        L.garage_bm_default_config.restype = None
        L.garage_bm_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.garage_bm_destroy.argtypes = [vp]
        L.garage_bm_destroy.restype = None
        L.garage_bm_blake2sum.argtypes = [vp, sz, vp]
        L.garage_bm_blake2sum.restype = None
        L.garage_bm_rpc_put_block.argtypes = [vp, vp, vp, sz]
        L.garage_bm_rpc_get_block.argtypes = [vp, vp, vp, sz, C.POINTER(sz)]
        L.garage_bm_resync_block.argtypes = [vp, i32, vp]
        L.garage_bm_block_incref.argtypes = [vp, vp]
        L.garage_bm_block_decref.argtypes = [vp, vp]
        L.garage_bm_get_block_rc.argtypes = [vp, vp]
        L.garage_bm_get_block_rc.restype = C.c_longlong
        L.garage_bm_resync_all.argtypes = [vp, i32, i32, C.POINTER(C.c_uint64)]
        L.garage_bm_repair_enqueue_missing.argtypes = [vp, i32, C.POINTER(C.c_uint64)]
        L.garage_bm_scrub.argtypes = [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.garage_bm_scrub_step.argtypes = [vp, i32, vp, sz, vp, C.POINTER(i32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.garage_bm_set_node_up.argtypes = [vp, i32, i32]
        L.garage_bm_corrupt_shard.argtypes = [vp, i32, vp, sz]
        L.garage_bm_drop_shard.argtypes = [vp, i32, vp]
        L.garage_bm_node_shard_index.argtypes = [vp, i32, vp]
        L.garage_bm_storage_nodes_of.argtypes = [vp, vp, C.POINTER(i32)]
        L.garage_bm_get_metrics.argtypes = [vp, C.POINTER(Metrics)]
        L.garage_bm_get_metrics.restype = None
        L.garage_bm_bench.argtypes = [vp, i32, i32, C.c_uint32, i32, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.garage_bm_corrupt_shard_header.argtypes = [vp, i32, vp, i32]
        L.garage_bm_plant_stale_shard.argtypes = [vp, i32, vp]
        L.garage_bm_set_node_readonly.argtypes = [vp, i32, i32]
        _lib = L
    return _lib


def blake2sum(data) -> bytes:
    L = load_library()
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data)
    out = np.zeros(32, dtype=np.uint8)
    L.garage_bm_blake2sum(C.c_void_p(a.ctypes.data) if a.size else None, a.size, C.c_void_p(out.ctypes.data))
    return out.tobytes()


class BlockManagerError(RuntimeError):
    def __init__(self, code):
        super().__init__("garage_bm error %d" % code)
        self.code = code


class BlockManager:
    """Same method names as garage_block::manager::BlockManager where they exist."""

    def __init__(self, k=10, m=4, n_nodes=None, cuda_device=0, batch_max_blocks=64, batch_linger_us=100,
                 block_ram_buffer_max=256 << 20, data_dir=None, shard_sum_kind=SUM_ADLER8, verify_content_hash=True,
                 data_fsync=False, block_gc_delay_ms=0):
        """block_gc_delay_ms defaults to 0 HERE (tests delete at once); the library's default is the
        reference's BLOCK_GC_DELAY of 10 minutes."""
        self._L = load_library()
        cfg = Config()
        self._L.garage_bm_default_config(C.byref(cfg))
        cfg.data_shards, cfg.parity_shards, cfg.cuda_device = k, m, cuda_device
        cfg.n_nodes = n_nodes or (k + m)
        cfg.batch_max_blocks, cfg.batch_linger_us = batch_max_blocks, batch_linger_us
        cfg.block_ram_buffer_max = block_ram_buffer_max
        cfg.shard_sum_kind, cfg.verify_content_hash = shard_sum_kind, 1 if verify_content_hash else 0
        cfg.data_fsync, cfg.block_gc_delay_ms = 1 if data_fsync else 0, block_gc_delay_ms
        self._dir = data_dir.encode() if data_dir else None  # keep the bytes alive
        cfg.data_dir = self._dir
        h = C.c_void_p()
        rc = self._L.garage_bm_create(C.byref(h), C.byref(cfg))
        if rc != OK:
            raise BlockManagerError(rc)
        self._h, self.k, self.m, self.n_nodes = h, k, m, cfg.n_nodes

    def close(self):
        if getattr(self, "_h", None):
            self._L.garage_bm_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _hp(hash32):
        return C.c_char_p(bytes(hash32))

    def rpc_put_block(self, hash32, data):
        a = np.ascontiguousarray(data, dtype=np.uint8)
        return self._L.garage_bm_rpc_put_block(self._h, self._hp(hash32), C.c_void_p(a.ctypes.data), a.size)

    def rpc_get_block(self, hash32, cap=(1 << 20) + 64):
        out = np.empty(cap, dtype=np.uint8)
        n = C.c_size_t(0)
        rc = self._L.garage_bm_rpc_get_block(self._h, self._hp(hash32), C.c_void_p(out.ctypes.data), cap, C.byref(n))
        return rc, (out[: n.value].copy() if rc == OK else None)

    def block_incref(self, hash32):
        return self._L.garage_bm_block_incref(self._h, self._hp(hash32))

    def block_decref(self, hash32):
        return self._L.garage_bm_block_decref(self._h, self._hp(hash32))

    def get_block_rc(self, hash32):
        return int(self._L.garage_bm_get_block_rc(self._h, self._hp(hash32)))

    def resync_block(self, node, hash32):
        return self._L.garage_bm_resync_block(self._h, node, self._hp(hash32))

    def resync_all(self, node, workers=8):
        done = C.c_uint64(0)
        failed = self._L.garage_bm_resync_all(self._h, node, workers, C.byref(done))
        return failed, done.value

    def repair_enqueue_missing(self, node):
        n = C.c_uint64(0)
        self._L.garage_bm_repair_enqueue_missing(self._h, node, C.byref(n))
        return n.value

    def scrub(self, node):
        a, b = C.c_uint64(0), C.c_uint64(0)
        rc = self._L.garage_bm_scrub(self._h, node, C.byref(a), C.byref(b))
        return rc, a.value, b.value

    def scrub_step(self, node, cursor=None, max_shards=0):
        """-> (rc, cursor_out bytes, finished, checked, corrupt)"""
        out = C.create_string_buffer(32)
        fin, a, b = C.c_int(0), C.c_uint64(0), C.c_uint64(0)
        rc = self._L.garage_bm_scrub_step(self._h, node, C.c_char_p(bytes(cursor)) if cursor else None, max_shards,
                                          C.cast(out, C.c_void_p), C.byref(fin), C.byref(a), C.byref(b))
        return rc, out.raw, bool(fin.value), a.value, b.value

    def set_node_up(self, node, up):
        return self._L.garage_bm_set_node_up(self._h, node, 1 if up else 0)

    def corrupt_shard(self, node, hash32, byte_off=0):
        return self._L.garage_bm_corrupt_shard(self._h, node, self._hp(hash32), byte_off)

    def corrupt_shard_header(self, node, hash32, what):
        return self._L.garage_bm_corrupt_shard_header(self._h, node, self._hp(hash32), what)

    def plant_stale_shard(self, node, hash32):
        return self._L.garage_bm_plant_stale_shard(self._h, node, self._hp(hash32))

    def set_node_readonly(self, node, readonly):
        return self._L.garage_bm_set_node_readonly(self._h, node, 1 if readonly else 0)

    def bench(self, threads, blocks_per_thread, block_len=1 << 20, mode=0, seed=1):
        """native closed-loop load generator: -> (rc, GiB/s, errors)"""
        g, e = C.c_double(0), C.c_uint64(0)
        rc = self._L.garage_bm_bench(self._h, threads, blocks_per_thread, block_len, mode, seed, C.byref(g), C.byref(e))
        return rc, g.value, e.value

    def drop_shard(self, node, hash32):
        return self._L.garage_bm_drop_shard(self._h, node, self._hp(hash32))

    def node_shard_index(self, node, hash32):
        return self._L.garage_bm_node_shard_index(self._h, node, self._hp(hash32))

    def storage_nodes_of(self, hash32):
        out = (C.c_int * (self.k + self.m))()
        self._L.garage_bm_storage_nodes_of(self._h, self._hp(hash32), out)
        return list(out)

    def metrics(self):
        mt = Metrics()
        self._L.garage_bm_get_metrics(self._h, C.byref(mt))
        return {n: getattr(mt, n) for n, _ in Metrics._fields_}


# ---- shard wire format (include/garage_shard_wire.h) ---------------------------------------------
class ShardHeader(C.Structure):
    _fields_ = [("hash", C.c_uint8 * 32), ("header", C.c_uint8), ("k", C.c_uint8), ("m", C.c_uint8),
                ("index", C.c_uint8), ("sum_kind", C.c_uint8), ("migrated", C.c_uint8),
                ("block_len", C.c_uint32), ("shard_len", C.c_uint32), ("sum", C.c_uint8 * 32)]

    def as_dict(self):
        return {"hash": bytes(self.hash), "header": int(self.header), "k": int(self.k), "m": int(self.m),
                "index": int(self.index), "sum_kind": int(self.sum_kind), "migrated": int(self.migrated),
                "block_len": int(self.block_len), "shard_len": int(self.shard_len), "sum": bytes(self.sum)}


WIRE_SYMBOLS = ["garage_shard_wire_encode", "garage_shard_wire_decode", "garage_shard_wire_encode_v0"]


def _wire():
    L = load_library()
    L.garage_shard_wire_encode.argtypes = [C.POINTER(ShardHeader), C.c_void_p, C.c_size_t]
    L.garage_shard_wire_encode.restype = C.c_size_t
    L.garage_shard_wire_decode.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(ShardHeader)]
    L.garage_shard_wire_encode_v0.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.garage_shard_wire_encode_v0.restype = C.c_size_t
    return L


def wire_encode(hash32, header, k, m, index, block_len, shard_len, sum_kind, sum32) -> bytes:
    h = ShardHeader()
    C.memmove(h.hash, bytes(hash32), 32)
    C.memmove(h.sum, bytes(sum32), 32)
    h.header, h.k, h.m, h.index, h.sum_kind = header, k, m, index, sum_kind
    h.block_len, h.shard_len = block_len, shard_len
    buf = (C.c_uint8 * 256)()
    n = _wire().garage_shard_wire_encode(C.byref(h), buf, 256)
    return bytes(buf[:n])


def wire_encode_v0(hash32, header) -> bytes:
    buf = (C.c_uint8 * 128)()
    hb = (C.c_uint8 * 32).from_buffer_copy(bytes(hash32))
    n = _wire().garage_shard_wire_encode_v0(hb, header, buf, 128)
    return bytes(buf[:n])


def wire_decode(data: bytes):
    h = ShardHeader()
    b = (C.c_uint8 * len(data)).from_buffer_copy(data) if data else None
    rc = _wire().garage_shard_wire_decode(b, len(data), C.byref(h))
    return None if rc else h.as_dict()
