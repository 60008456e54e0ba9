"""Wire / metadata encoding of one erasure-coded shard (rows a8 / f3, include/garage_shard_wire.h):
msgpack struct map behind a version marker, decoded the way garage_util::migrate::Migrate::decode
works (src/util/migrate.rs:19-29): current format first, else the previous (replicated PutBlock,
src/block/manager.rs:54-69) migrated.  PINNED by the independent `msgpack` python package: our bytes
are exactly what msgpack.packb produces for the same map, and msgpack.unpackb reads them back."""
import hashlib
import os
import sys

import msgpack
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from garage_b200 import block_manager as BM  # noqa: E402

H = hashlib.blake2b(b"block").digest()[:32]
S = hashlib.blake2b(b"shard").digest()[:32]


def want_v1(header, k, m, index, block_len, shard_len, sum_kind):
    d = {"hash": H, "header": "Compressed" if header else "Plain", "k": k, "m": m, "index": index,
         "block_len": block_len, "shard_len": shard_len, "sum_kind": sum_kind, "sum": S}
    return b"GEC1shdr" + msgpack.packb(d, use_bin_type=True)


@pytest.mark.parametrize("args", [(0, 10, 4, 13, 1 << 20, 104858, 1), (1, 6, 3, 0, 777777, 129630, 0),
                                  (0, 4, 2, 5, 1, 1, 1), (0, 32, 8, 39, 0xFFFFFFFF, 134217728, 0),
                                  (0, 1, 1, 1, 200, 200, 0), (1, 16, 4, 7, 65536, 4096, 1)])
def test_v1_bytes_equal_msgpack_and_round_trip(args):
    header, k, m, index, block_len, shard_len, sum_kind = args
    enc = BM.wire_encode(H, header, k, m, index, block_len, shard_len, sum_kind, S)
    assert enc == want_v1(*args)
    back = msgpack.unpackb(enc[8:], raw=False)
    assert back["hash"] == H and back["sum"] == S and back["index"] == index and back["block_len"] == block_len
    d = BM.wire_decode(enc)
    assert d == {"hash": H, "header": header, "k": k, "m": m, "index": index, "sum_kind": sum_kind, "migrated": 0,
                 "block_len": block_len, "shard_len": shard_len, "sum": S}


def test_previous_format_is_migrated_like_migrate_decode():
    """what a not-yet-upgraded node sends: rmp-serde struct map of PutBlock{hash, header}"""
    for header, name in ((0, "Plain"), (1, "Compressed")):
        v0 = msgpack.packb({"hash": H, "header": name}, use_bin_type=True)
        assert BM.wire_encode_v0(H, header) == v0
        d = BM.wire_decode(v0)
        assert d["migrated"] == 1 and (d["k"], d["m"], d["index"]) == (1, 0, 0)
        assert d["hash"] == H and d["header"] == header and d["block_len"] == 0 and d["sum"] == bytes(32)


def test_rejects_garbage_truncation_and_bad_geometry():
    enc = BM.wire_encode(H, 0, 10, 4, 3, 1 << 20, 104858, 1, S)
    assert BM.wire_decode(b"") is None
    for cut in (1, 7, 8, 9, 40, len(enc) - 1):
        assert BM.wire_decode(enc[:cut]) is None, cut
    assert BM.wire_decode(enc + b"\x00") is None                      # trailing bytes
    assert BM.wire_decode(b"GEC9shdr" + enc[8:]) is None              # unknown version marker
    bad = bytearray(enc)
    bad[8 + 1 + 5 + 2 + 5] ^= 0xFF  # inside the hash: still decodes, different hash
    assert BM.wire_decode(bytes(bad))["hash"] != H
    # index outside the code is refused on both sides
    assert BM.wire_encode(H, 0, 10, 4, 14, 1, 1, 0, S) == b""
    forged = b"GEC1shdr" + msgpack.packb({"hash": H, "header": "Plain", "k": 4, "m": 2, "index": 6, "block_len": 1,
                                          "shard_len": 1, "sum_kind": 0, "sum": S}, use_bin_type=True)
    assert BM.wire_decode(forged) is None
    unknown = b"GEC1shdr" + msgpack.packb({"hash": H, "header": "Plain", "k": 4, "m": 2, "index": 1, "block_len": 1,
                                           "shard_len": 1, "sum_kind": 0, "sum": S, "extra": 1}, use_bin_type=True)
    assert BM.wire_decode(unknown) is None


def test_library_exports_wire_symbols():
    L = BM.load_library()
    for name in BM.WIRE_SYMBOLS:
        assert hasattr(L, name), name
