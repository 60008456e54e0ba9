"""The C++ mirror of BlockManager (libgarage_block.so) driven the way the reference's own tests
drive the block path: PUT objects of several sizes, GET them back and compare bytes
(src/garage/tests/s3/multipart.rs, script/test-smoke.sh:51-58), plus what erasure coding adds:
lost nodes, corrupt shards, resync of a wiped node, scrub."""
import hashlib
import os
import re
import subprocess
import sys
import threading

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402

from garage_b200 import block_manager as BM  # noqa: E402


# ---------------------------------------------------------------- CPU-only boundary checks
def test_library_exports_every_declared_symbol():
    BM.load_library()
    src = open(os.path.join(ROOT, "include", "garage_block_manager.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decl = sorted(set(re.findall(r"\b(garage_bm_[a-z0-9_]+)\s*\(", src)))
    assert decl == sorted(BM.SYMBOLS)
    from garage_b200 import _build

    out = subprocess.run(["nm", "-D", "--defined-only", _build.BM_SO], capture_output=True, text=True).stdout
    assert set(decl) <= set(re.findall(r"\bT (garage_bm_[a-z0-9_]+)", out))
    # the host mirror reaches the GPU only through the C ABI: no CUDA runtime symbols of its own
    und = subprocess.run(["nm", "-D", "--undefined-only", _build.BM_SO], capture_output=True, text=True).stdout
    assert "garage_ec_encode_blocks_with_sums" in und and "cudaMalloc" not in und and "cudaLaunch" not in und


def test_blake2sum_is_garages_content_hash():
    for n in (0, 1, 3072, 3073, 1 << 20):
        b = O.fill_random(n, n + 1)
        assert BM.blake2sum(b) == hashlib.blake2b(b.tobytes()).digest()[:32]  # util/data.rs:130-138


def _no_cuda():
    import torch

    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_cuda(), reason="checks the no-GPU behaviour")
def test_no_gpu_no_block_manager():
    with pytest.raises(BM.BlockManagerError) as e:
        BM.BlockManager(4, 2)
    assert e.value.code == -5  # GARAGE_EC_E_NODEVICE: no CPU fallback


# ---------------------------------------------------------------- GPU: behaviour
SIZES = [3073, 65536, 1 << 20, (1 << 20) - 1, 5, 777777, 1048575, 4096]  # > INLINE_THRESHOLD (manager.rs:46) except 5


def put_all(bm, blocks):
    hashes = []
    for b in blocks:
        h = BM.blake2sum(b)  # put.rs:448
        assert bm.rpc_put_block(h, b) == BM.OK
        hashes.append(h)
    return hashes


@pytest.mark.gpu
@pytest.mark.parametrize("k,m", [(4, 2), (10, 4), (6, 3)])
def test_put_get_roundtrip(k, m):
    blocks = [O.fill_random(n, 100 + i) for i, n in enumerate(SIZES)]
    with BM.BlockManager(k, m) as bm:
        hashes = put_all(bm, blocks)
        for h, b in zip(hashes, blocks):
            rc, got = bm.rpc_get_block(h)
            assert rc == BM.OK and np.array_equal(got, b)
            who = bm.storage_nodes_of(h)
            assert sorted(who) == list(range(k + m))
            assert [bm.node_shard_index(who[i], h) for i in range(k + m)] == list(range(k + m))
        assert bm.metrics()["reconstruct_calls"] == 0  # all data shards present: no GPU on GET
        rc, _ = bm.rpc_get_block(BM.blake2sum(b"never stored"))
        assert rc == BM.E_MISSING_BLOCK


@pytest.mark.gpu
def test_get_survives_m_lost_nodes_and_reports_missing_beyond():
    k, m = 10, 4
    blocks = [O.fill_random(n, 7 + i) for i, n in enumerate(SIZES)]
    with BM.BlockManager(k, m, n_nodes=16) as bm:
        hashes = put_all(bm, blocks)
        for dead in ([0, 1, 2, 3], [15, 7, 9, 12], [5]):
            for d in dead:
                bm.set_node_up(d, False)
            for h, b in zip(hashes, blocks):
                rc, got = bm.rpc_get_block(h)
                assert rc == BM.OK and np.array_equal(got, b), dead
            for d in dead:
                bm.set_node_up(d, True)
        assert bm.metrics()["reconstruct_calls"] > 0
        for d in range(5):
            bm.set_node_up(d, False)
        lost = 0
        for h in hashes:
            who = bm.storage_nodes_of(h)
            down = sum(1 for w in who if w < 5)
            rc, _ = bm.rpc_get_block(h)
            assert rc == (BM.E_MISSING_BLOCK if down > m else BM.OK)
            lost += rc != BM.OK
        assert lost > 0


@pytest.mark.gpu
def test_put_quorum():
    k, m = 4, 2
    b = O.fill_random(200000, 1)
    with BM.BlockManager(k, m) as bm:
        bm.set_node_up(0, False)
        assert bm.rpc_put_block(BM.blake2sum(b), b) == BM.OK  # 5 of 6 stored >= k+1
        bm.set_node_up(1, False)
        b2 = O.fill_random(200000, 2)
        assert bm.rpc_put_block(BM.blake2sum(b2), b2) == BM.E_QUORUM  # 4 < k+1


@pytest.mark.gpu
def test_corrupt_shard_is_quarantined_then_resynced():
    k, m = 6, 3
    b = O.fill_random(1 << 20, 3)
    with BM.BlockManager(k, m) as bm:
        h = BM.blake2sum(b)
        assert bm.rpc_put_block(h, b) == BM.OK
        who = bm.storage_nodes_of(h)
        assert bm.corrupt_shard(who[2], h, 12345) == BM.OK   # a data shard
        assert bm.corrupt_shard(who[7], h, 1) == BM.OK       # a parity shard (not read while data is complete)
        rc, got = bm.rpc_get_block(h)                        # read_block_from: mismatch -> quarantine + resync queue
        assert rc == BM.OK and np.array_equal(got, b)
        mt = bm.metrics()
        assert mt["corruption_counter"] == 1 and mt["resync_queue_length"] == 1
        assert bm.node_shard_index(who[2], h) == -1
        failed, done = bm.resync_all(who[2])
        assert (failed, done) == (0, 1) and bm.node_shard_index(who[2], h) == 2
        # scrub finds the parity corruption the GET never touched
        rc, checked, corrupt = bm.scrub(who[7])
        assert rc == BM.OK and checked == 1 and corrupt == 1
        assert bm.resync_all(who[7]) == (0, 1)
        for n in who:
            rc, checked, corrupt = bm.scrub(n)
            assert (rc, corrupt) == (BM.OK, 0)
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.OK and np.array_equal(got, b)


@pytest.mark.gpu
def test_wiped_node_is_rebuilt_by_repair_and_resync_workers():
    """`garage repair blocks` on a replaced node (doc/book/operations/durability-repairs.md): every
    shard it should hold is re-created from k survivors; 8 workers share the GPU through the batcher."""
    k, m, nblocks = 10, 4, 96
    rng = np.random.default_rng(0)
    blocks = [O.fill_random(int(rng.integers(3073, (1 << 20) + 1)), 50 + i) for i in range(nblocks)]
    with BM.BlockManager(k, m, batch_max_blocks=32, batch_linger_us=2000) as bm:
        hashes = put_all(bm, blocks)
        victim = 3
        before = {h: bm.node_shard_index(victim, h) for h in hashes}
        for h in hashes:
            assert bm.drop_shard(victim, h) == BM.OK
        assert bm.repair_enqueue_missing(victim) == nblocks
        failed, done = bm.resync_all(victim, workers=8)
        assert (failed, done) == (0, nblocks)
        mt = bm.metrics()
        assert mt["resync_counter"] == nblocks and mt["reconstruct_calls"] == nblocks
        assert mt["reconstruct_batches"] < nblocks  # calls were coalesced into GPU batches (row f1)
        assert {h: bm.node_shard_index(victim, h) for h in hashes} == before
        rc, checked, corrupt = bm.scrub(victim)      # rebuilt shards carry correct per-shard sums
        assert (rc, checked, corrupt) == (BM.OK, nblocks, 0)
        # and they are the right bytes: lose m OTHER nodes, so the rebuilt shard is needed to decode
        for d in (0, 1, 2, 4):
            bm.set_node_up(d, False)
        for h, b in zip(hashes, blocks):
            rc, got = bm.rpc_get_block(h)
            assert rc == BM.OK and np.array_equal(got, b)


@pytest.mark.gpu
def test_concurrent_puts_are_batched():
    k, m, per, nthreads = 10, 4, 12, 8
    blocks = [[O.fill_random(1 << 20, 1000 + t * per + i) for i in range(per)] for t in range(nthreads)]
    with BM.BlockManager(k, m, batch_max_blocks=64, batch_linger_us=3000) as bm:
        errs = []

        def worker(t):
            for b in blocks[t]:
                if bm.rpc_put_block(BM.blake2sum(b), b) != BM.OK:
                    errs.append(t)

        th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
        [x.start() for x in th]
        [x.join() for x in th]
        assert not errs
        mt = bm.metrics()
        assert mt["put_calls"] == per * nthreads and mt["put_batches"] < mt["put_calls"]
        for t in range(nthreads):
            for b in blocks[t]:
                rc, got = bm.rpc_get_block(BM.blake2sum(b))
                assert rc == BM.OK and np.array_equal(got, b)


@pytest.mark.gpu
def test_file_backed_store_survives_restart(tmp_path):
    """row f3: shards on disk as <data_dir>/node<N>/<hh>/<hh>/<hash>.shard (header + bytes), written
    tmp -> rename; a new BlockManager on the same directory serves the same blocks; a corrupt file
    is renamed *.corrupted and rebuilt by resync."""
    k, m = 4, 2
    blocks = [O.fill_random(n, 300 + i) for i, n in enumerate([1 << 20, 3073, 500000])]
    d = str(tmp_path / "data")
    with BM.BlockManager(k, m, data_dir=d) as bm:
        hashes = put_all(bm, blocks)
    files = sorted(p for p in (tmp_path / "data").rglob("*.shard"))
    assert len(files) == len(blocks) * (k + m)
    f0 = files[0]
    assert f0.parent.name == f0.name[2:4] and f0.parent.parent.name == f0.name[0:2]  # layout.rs:286-291
    raw = f0.read_bytes()
    assert raw[:4] == b"GEC1" and raw[4] == k and raw[5] == m
    shard_len = int.from_bytes(raw[12:16], "little")
    assert len(raw) == 64 + shard_len
    # per-shard tag in the header, of the kind byte 7 names (default: adler8 = 8 x zlib Adler-32)
    import struct
    import zlib

    assert raw[7] == BM.SUM_ADLER8
    body = raw[64:]
    seg = (((len(body) + 7) // 8) + 15) // 16 * 16
    assert raw[16:48] == b"".join(struct.pack("<I", zlib.adler32(body[i * seg:(i + 1) * seg])) for i in range(8))
    assert int.from_bytes(raw[48:52], "little") == zlib.adler32(raw[:48])  # header check covers every field
    assert not list((tmp_path / "data").rglob("*.tmp*"))
    with BM.BlockManager(k, m, data_dir=d) as bm:  # "restart"
        for h, b in zip(hashes, blocks):
            rc, got = bm.rpc_get_block(h)
            assert rc == BM.OK and np.array_equal(got, b)
        who = bm.storage_nodes_of(hashes[0])
        assert bm.corrupt_shard(who[1], hashes[0], 77) == BM.OK
        rc, got = bm.rpc_get_block(hashes[0])
        assert rc == BM.OK and np.array_equal(got, blocks[0])
        assert len(list((tmp_path / "data").rglob("*.corrupted"))) == 1
        assert bm.resync_all(who[1]) == (0, 1)
        rc, checked, corrupt = bm.scrub(who[1])
        assert (rc, corrupt) == (BM.OK, 0) and checked >= 1


@pytest.mark.gpu
def test_refcount_drives_resync_and_deletion():
    """block_incref / block_decref (manager.rs:452-500): 0 -> 1 queues a safety resync everywhere,
    the last decref queues the deletion; resync then deletes the shard (delete_if_unneeded)."""
    k, m = 4, 2
    b = O.fill_random(300000, 9)
    with BM.BlockManager(k, m) as bm:
        h = BM.blake2sum(b)
        assert bm.get_block_rc(h) == -1
        bm.block_incref(h)                      # BlockRefTable::updated on insert
        assert bm.get_block_rc(h) == 1 and bm.metrics()["resync_queue_length"] == k + m
        assert bm.rpc_put_block(h, b) == BM.OK
        for node in range(k + m):                # the safety resync finds every shard in place
            assert bm.resync_all(node) == (0, 1)
        bm.block_incref(h)
        bm.block_decref(h)
        assert bm.get_block_rc(h) == 1 and bm.metrics()["resync_queue_length"] == 0
        # a lost shard of a referenced block is rebuilt ...
        who = bm.storage_nodes_of(h)
        assert bm.drop_shard(who[0], h) == BM.OK
        assert bm.resync_block(who[0], h) == BM.OK and bm.node_shard_index(who[0], h) == 0
        # ... and after the last decref every node deletes its shard
        bm.block_decref(h)
        assert bm.get_block_rc(h) == 0 and bm.metrics()["resync_queue_length"] == k + m
        for node in range(k + m):
            assert bm.resync_all(node) == (0, 1)
            assert bm.node_shard_index(node, h) == -1
        rc, _ = bm.rpc_get_block(h)
        assert rc == BM.E_MISSING_BLOCK
        assert bm.metrics()["delete_counter"] == k + m + 1  # + the drop_shard above


@pytest.mark.gpu
def test_scrub_is_resumable_from_a_checkpoint():
    """ScrubWorker checkpoints its iterator (repair.rs:186-193,460-464): scrub_step visits shards in
    hash order from a cursor, so a sweep can be cut in pieces (or survive a restart) and still see every
    shard exactly once."""
    k, m, nblocks = 4, 2, 23
    blocks = [O.fill_random(50000 + 1000 * i, 400 + i) for i in range(nblocks)]
    with BM.BlockManager(k, m) as bm:
        hashes = put_all(bm, blocks)
        node = 2
        bm.corrupt_shard(node, hashes[5], 3)
        bm.corrupt_shard(node, hashes[17], 9)
        cursor, total, bad, steps = None, 0, 0, 0
        while True:
            rc, cursor, finished, checked, corrupt = bm.scrub_step(node, cursor, max_shards=7)
            assert rc == BM.OK and checked <= 7
            total += checked
            bad += corrupt
            steps += 1
            if finished:
                break
        assert (total, bad, steps) == (nblocks, 2, 4)
        assert bm.metrics()["resync_queue_length"] == 2
        rc, cursor2, finished, checked, corrupt = bm.scrub_step(node, cursor, max_shards=7)
        assert finished and checked == 0 and cursor2 == cursor  # nothing after the last checkpoint


@pytest.mark.gpu
def test_blake2_shard_tags_still_supported(tmp_path):
    k, m = 4, 2
    b = O.fill_random(400000, 21)
    d = str(tmp_path / "data")
    with BM.BlockManager(k, m, data_dir=d, shard_sum_kind=BM.SUM_BLAKE2) as bm:
        h = BM.blake2sum(b)
        assert bm.rpc_put_block(h, b) == BM.OK
        raw = sorted((tmp_path / "data").rglob("*.shard"))[0].read_bytes()
        assert raw[7] == BM.SUM_BLAKE2 and raw[16:48] == hashlib.blake2b(raw[64:]).digest()[:32]
        rc, checked, corrupt = bm.scrub(bm.storage_nodes_of(h)[0])
        assert (rc, checked, corrupt) == (BM.OK, 1, 0)
    # a manager configured for adler8 reads (and scrubs, on the CPU) shards written with blake2 tags
    with BM.BlockManager(k, m, data_dir=d) as bm:
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.OK and np.array_equal(got, b)
        rc, checked, corrupt = bm.scrub(bm.storage_nodes_of(h)[1])
        assert (rc, checked, corrupt) == (BM.OK, 1, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("in_memory", [True, False])
def test_get_checks_the_content_hash_and_hunts_the_bad_shard(tmp_path, in_memory):
    """DataBlock::verify on every read (block.rs:69-83): a shard that passes its OWN tag but holds other
    bytes (stale / written by a bad resync) must not reach the client.  GET notices through the
    whole-block blake2sum, finds the culprit by leave-one-out reconstruction, quarantines it, and still
    returns the right block; resync then rebuilds it."""
    k, m = 6, 3
    b = O.fill_random(900001, 33)
    kw = {} if in_memory else {"data_dir": str(tmp_path / "d")}
    with BM.BlockManager(k, m, **kw) as bm:
        h = BM.blake2sum(b)
        assert bm.rpc_put_block(h, b) == BM.OK
        who = bm.storage_nodes_of(h)
        assert bm.plant_stale_shard(who[3], h) == BM.OK  # data shard 3: valid tag, wrong bytes
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.OK and np.array_equal(got, b)
        mt = bm.metrics()
        assert mt["corrupt_data_errors"] == 1 and mt["corruption_counter"] == 1 and mt["resync_queue_length"] == 1
        assert bm.node_shard_index(who[3], h) == -1  # quarantined
        assert bm.resync_all(who[3]) == (0, 1)
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.OK and np.array_equal(got, b) and bm.metrics()["corrupt_data_errors"] == 1
        # more stale shards than the code can route around: CorruptData, never wrong bytes
        for i in (0, 1, 2, 4):
            assert bm.plant_stale_shard(who[i], h) == BM.OK
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.E_CORRUPT_DATA and got is None


@pytest.mark.gpu
@pytest.mark.parametrize("in_memory", [True, False])
def test_corrupt_shard_metadata_is_caught(tmp_path, in_memory):
    """flipped block_len / index in a shard's header: covered by the header check (file store) and
    cross-checked against the other shards, never trusted"""
    k, m = 4, 2
    b = O.fill_random(600000, 44)
    kw = {} if in_memory else {"data_dir": str(tmp_path / "d")}
    with BM.BlockManager(k, m, **kw) as bm:
        h = BM.blake2sum(b)
        assert bm.rpc_put_block(h, b) == BM.OK
        who = bm.storage_nodes_of(h)
        assert bm.corrupt_shard_header(who[0], h, 1) == BM.OK  # block_len + 1
        assert bm.corrupt_shard_header(who[2], h, 2) == BM.OK  # index + 1
        rc, got = bm.rpc_get_block(h)
        assert rc == BM.OK and np.array_equal(got, b)
        assert bm.metrics()["corruption_counter"] == 2 and bm.metrics()["resync_queue_length"] == 2
        assert bm.resync_all(who[0]) == (0, 1) and bm.resync_all(who[2]) == (0, 1)
        for n in who:
            assert bm.scrub(n)[2] == 0


@pytest.mark.gpu
def test_failed_shard_writes_do_not_count_towards_the_quorum(tmp_path):
    k, m = 4, 2
    with BM.BlockManager(k, m, data_dir=str(tmp_path / "d"), data_fsync=True) as bm:
        b = O.fill_random(200000, 5)
        h = BM.blake2sum(b)
        who = bm.storage_nodes_of(h)
        assert bm.rpc_put_block(h, b) == BM.OK and bm.metrics()["write_errors"] == 0
        b2 = O.fill_random(200000, 6)
        h2 = BM.blake2sum(b2)
        who2 = bm.storage_nodes_of(h2)
        for n in who2[:2]:
            assert bm.set_node_readonly(n, True) == BM.OK
        assert bm.rpc_put_block(h2, b2) == BM.E_QUORUM  # 4 stored < k + 1
        assert bm.metrics()["write_errors"] == 2
        for n in who2[:2]:
            bm.set_node_readonly(n, False)


@pytest.mark.gpu
def test_block_gc_delay_protects_a_block_that_is_referenced_again():
    """BLOCK_GC_DELAY (manager.rs:49-52): a shard whose block dropped to rc 0 is only deleted after the
    delay, and the decision is re-checked under the locks -- an incref in between keeps the shard."""
    k, m = 4, 2
    b = O.fill_random(123456, 8)
    with BM.BlockManager(k, m, block_gc_delay_ms=60000) as bm:
        h = BM.blake2sum(b)
        bm.block_incref(h)
        assert bm.rpc_put_block(h, b) == BM.OK
        for node in range(k + m):
            bm.resync_all(node)
        bm.block_decref(h)  # rc 0: deletion queued ...
        for node in range(k + m):
            assert bm.resync_all(node) == (0, 1)
            assert bm.node_shard_index(node, h) >= 0  # ... but inside the GC delay nothing is deleted
        assert bm.metrics()["delete_counter"] == 0 and bm.metrics()["resync_queue_length"] == k + m
        bm.block_incref(h)  # referenced again (re-upload of the same content)
        for node in range(k + m):
            bm.resync_all(node)
            assert bm.node_shard_index(node, h) >= 0
        assert bm.metrics()["delete_counter"] == 0


@pytest.mark.gpu
def test_native_load_generator_round_trip():
    with BM.BlockManager(10, 4, batch_max_blocks=32, batch_linger_us=500) as bm:
        rc, gibs, errs = bm.bench(8, 6, 1 << 20, 0, seed=3)
        assert (rc, errs) == (0, 0) and gibs > 0
        rc, gibs, errs = bm.bench(8, 6, 1 << 20, 1, seed=3)
        assert (rc, errs) == (0, 0)
        for d in range(4):
            bm.set_node_up(d, False)
        rc, gibs, errs = bm.bench(8, 6, 1 << 20, 1, seed=3)
        assert (rc, errs) == (0, 0) and bm.metrics()["reconstruct_calls"] > 0
