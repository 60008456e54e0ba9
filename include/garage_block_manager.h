/*
 * garage_block_manager.h -- C view of the C++ host-side mirror of Garage's
 * garage_block::manager::BlockManager for the erasure-coded block path
 * (garage_b200/csrc/block_manager.{h,cpp} -> libgarage_block.so).
 *
 * Garage is Rust and there is no Rust toolchain in this environment, so the host side above
 * the C ABI (include/garage_ec.h) is written in C++ with the reference's names, argument
 * meaning and error behaviour; it talks to the GPU ONLY through garage_ec_* (it is exactly the
 * code a Rust maintainer would write in src/block).  A "cluster" is
 * n_nodes in-process node stores (the reference tests multi-node the same way: several
 * instances on localhost in one process, src/net/test.rs:15-60, script/test-smoke.sh).
 *
 * Mirrored reference surface:
 *   rpc_put_block           src/block/manager.rs:366-408   (encode, shard i -> node who[i])
 *   rpc_get_block           src/block/manager.rs:344-363, 276-339 (any k of k+m, reconstruct)
 *   resync_block (fetch)    src/block/resync.rs:460-500    (rebuild this node's shard)
 *   read_block / verify     src/block/manager.rs:554-609, src/block/block.rs:69-83
 *                           (per-shard blake2sum; corrupt -> quarantine + resync queue)
 *   ScrubWorker             src/block/repair.rs:438-490    (sweep of one node's shards)
 *   write_block             src/block/manager.rs:517-530, 720-805 (local durable store)
 *   block_incref/decref     src/block/manager.rs:452-500, get_block_rc :421-449 (src/block/rc.rs)
 *   batching front-end      (row f1) concurrent PUT / resync calls coalesced into GPU batches,
 *                           back-pressure like buffer_kb_semaphore (manager.rs:156,380-385)
 */
#ifndef GARAGE_BLOCK_MANAGER_H
#define GARAGE_BLOCK_MANAGER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* errors: mirror garage_util::error::Error variants on this path (src/util/error.rs:14-82) */
#define GARAGE_BM_OK 0
#define GARAGE_BM_E_CORRUPT_DATA (-10)  /* Error::CorruptData(hash)  */
#define GARAGE_BM_E_MISSING_BLOCK (-11) /* Error::MissingBlock(hash) */
#define GARAGE_BM_E_MESSAGE (-12)       /* Error::Message(..)        */
#define GARAGE_BM_E_QUORUM (-13)        /* Error::Quorum(..) from try_write_many_sets */
/* GARAGE_EC_E_* codes (-1..-6) pass through unchanged */

typedef struct garage_bm garage_bm;

typedef struct {
    int data_shards;                /* k  (config: erasure_coding.data_shards)   */
    int parity_shards;              /* m                                          */
    int cuda_device;
    int n_nodes;                    /* >= k+m storage nodes in the simulated cluster */
    uint32_t block_size;            /* util/config.rs:273-275, default 1 MiB          */
    uint64_t block_ram_buffer_max;  /* util/config.rs:276-278, default 256 MiB        */
    uint32_t batch_max_blocks;      /* f1: max blocks per GPU batch                   */
    uint32_t batch_linger_us;       /* f1: how long the first block of a batch waits  */
    const char *data_dir;           /* NULL/"": shards in memory; else <data_dir>/node<N>/<hh>/<hh>/<hash>.shard
                                       (64-byte header {GEC1,k,m,index,sum_kind,block_len,shard_len,tag,header check}
                                       + bytes; tmp-file -> rename, *.corrupted quarantine: manager.rs:720-819)  */
    int shard_sum_kind;             /* GARAGE_EC_SUM_ADLER8 (default) or GARAGE_EC_SUM_BLAKE2: the per-shard tag  */
    int verify_content_hash;        /* 1 (default): rpc_get_block checks blake2sum(block) == hash like
                                       DataBlock::verify (block.rs:69-83) and hunts the bad shard on mismatch     */
    int data_fsync;                 /* util/config.rs data_fsync: fsync file + directory before/after the rename  */
    uint32_t block_gc_delay_ms;     /* BLOCK_GC_DELAY (manager.rs:49-52, 10 min): a shard whose block dropped to
                                       rc 0 is deleted by resync only after this delay                            */
} garage_bm_config;

typedef struct {
    uint64_t bytes_written, bytes_read;      /* block.bytes_written / bytes_read  (metrics.rs) */
    uint64_t corruption_counter;             /* block.corruption_counter                        */
    uint64_t resync_counter, resync_error_counter, resync_recv_counter;
    uint64_t delete_counter;
    uint64_t put_calls, put_batches;         /* f1: how well PUTs were coalesced                */
    uint64_t reconstruct_calls, reconstruct_batches;
    uint64_t scrub_shards_checked, scrub_corruptions;
    uint64_t resync_queue_length;            /* block.resync_queue_length (all nodes)            */
    uint64_t encode_call_us, reconstruct_call_us; /* wall time spent inside garage_ec_* batch calls */
    uint64_t corrupt_data_errors;            /* GETs whose reassembled block failed the content hash */
    uint64_t write_errors;                   /* shard writes that failed (ENOSPC, EACCES, ...)      */
} garage_bm_metrics;

void garage_bm_default_config(garage_bm_config *cfg);
int garage_bm_create(garage_bm **out, const garage_bm_config *cfg);
void garage_bm_destroy(garage_bm *bm);

/* util/data.rs:130-138 */
void garage_bm_blake2sum(const uint8_t *data, size_t len, uint8_t hash_out[32]);

/* manager.rs:366-408.  hash = blake2sum(data) (put.rs:448).  Thread-safe; concurrent calls are
 * batched.  Returns GARAGE_BM_E_QUORUM if fewer than k+1 (or k+m if smaller) shards were stored. */
int garage_bm_rpc_put_block(garage_bm *bm, const uint8_t hash[32], const uint8_t *data, size_t len);

/* manager.rs:344-363: the block's bytes (whole buffer instead of a ByteStream).
 * GARAGE_BM_E_MISSING_BLOCK if fewer than k valid shards can be gathered.                       */
int garage_bm_rpc_get_block(garage_bm *bm, const uint8_t hash[32], uint8_t *out, size_t cap,
                            size_t *out_len);

/* manager.rs:452-500 block_incref / block_decref (called by BlockRefTable::updated,
 * src/model/s3/block_ref_table.rs:69-85): 0 -> 1 and 1 -> 0 transitions queue a resync on every
 * storage node of the block; get_block_rc (manager.rs:421-449) returns -1 for a block that was
 * never counted (treated as referenced).                                                        */
int garage_bm_block_incref(garage_bm *bm, const uint8_t hash[32]);
int garage_bm_block_decref(garage_bm *bm, const uint8_t hash[32]);
long long garage_bm_get_block_rc(garage_bm *bm, const uint8_t hash[32]);

/* resync.rs:354-503 for storage node `node`: fetch branch (rc > 0, shard absent: rebuild it from k
 * survivors) and the deletion half of the offload branch (rc == 0, shard present: delete it).   */
int garage_bm_resync_block(garage_bm *bm, int node, const uint8_t hash[32]);
/* drain node's resync queue with `workers` concurrent workers (resync.rs:43: up to 8);
 * returns the number of blocks that could not be resynced (they stay queued, with backoff
 * semantics left to the caller).                                                                */
int garage_bm_resync_all(garage_bm *bm, int node, int workers, uint64_t *resynced);
/* RepairWorker (repair.rs:35-155): enqueue every hash this node should hold but does not.        */
int garage_bm_repair_enqueue_missing(garage_bm *bm, int node, uint64_t *enqueued);

/* ScrubWorker sweep over all shards stored on `node` (repair.rs:438-490): GPU blake2sum of each
 * shard; a mismatch quarantines the shard (.corrupted, manager.rs:807-819) and queues a resync
 * (manager.rs:592-605).                                                                         */
int garage_bm_scrub(garage_bm *bm, int node, uint64_t *checked, uint64_t *corrupt);
/* the same sweep in resumable steps -- the scrub checkpoint of the reference (ScrubWorker persists its
 * BlockStoreIterator position every 60 s, repair.rs:186-193,460-464): shards are visited in hash order
 * starting after cursor32 (NULL = from the start), at most max_shards (0 = all); cursor_out32 is the
 * position to persist, *finished = 1 once the end of the node's store was reached.                */
int garage_bm_scrub_step(garage_bm *bm, int node, const uint8_t *cursor32, size_t max_shards,
                         uint8_t cursor_out32[32], int *finished, uint64_t *checked, uint64_t *corrupt);

/* native closed-loop load generator (row f1: the batching front-end decides real-world throughput):
 * `threads` client threads each PUT (mode 0) or GET (mode 1) `blocks_per_thread` blocks of `block_len`
 * bytes (splitmix64 stream of `seed`; a GET run must follow a PUT run with the same arguments).  The
 * content hashes are computed before the clock starts (put.rs:448 does that in the API layer).        */
int garage_bm_bench(garage_bm *bm, int threads, int blocks_per_thread, uint32_t block_len, int mode,
                    uint64_t seed, double *gib_per_s, uint64_t *errors);

/* fault injection / inspection for tests */
int garage_bm_set_node_up(garage_bm *bm, int node, int up);
int garage_bm_corrupt_shard(garage_bm *bm, int node, const uint8_t hash[32], size_t byte_off);
/* what = 1: stored block_len + 1; what = 2: stored shard index + 1 (header check left alone)          */
int garage_bm_corrupt_shard_header(garage_bm *bm, int node, const uint8_t hash[32], int what);
/* overwrite the shard with a self-consistent shard of OTHER content (valid tag, right index): only the
 * whole-block content hash can notice                                                                 */
int garage_bm_plant_stale_shard(garage_bm *bm, int node, const uint8_t hash[32]);
/* file store only: make every write to the node's directory fail (like ENOSPC / EACCES / a lost mount) */
int garage_bm_set_node_readonly(garage_bm *bm, int node, int readonly);
int garage_bm_drop_shard(garage_bm *bm, int node, const uint8_t hash[32]);
int garage_bm_node_shard_index(garage_bm *bm, int node, const uint8_t hash[32]); /* -1 if none */
int garage_bm_storage_nodes_of(garage_bm *bm, const uint8_t hash[32], int *nodes_out /* k+m */);
void garage_bm_get_metrics(garage_bm *bm, garage_bm_metrics *out);

#ifdef __cplusplus
}
#endif
#endif
