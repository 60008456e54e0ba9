// block_manager.cpp -- C++ host-side mirror of garage_block::manager::BlockManager for the
// erasure-coded block path (see include/garage_block_manager.h for the mirrored surface and
// the reference file:line of every method).  Host code only: it reaches the GPU exclusively
// through the C ABI in include/garage_ec.h, exactly like the Rust shim of INTEGRATION.md would.
//
// Row f1 (batching front-end): rpc_put_block / reconstructing GETs / resync workers are called
// from many threads (reference: <= 3 blocks in flight per PUT, src/api/s3/put.rs:42; 8 resync
// workers, src/block/resync.rs:43).  One block per GPU call cannot amortise launch + PCIe
// latency, so calls are queued and a dispatcher thread hands the GPU whole batches.
#include "../../include/garage_block_manager.h"
#include "../../include/garage_ec.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <cstdio>
#include <filesystem>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <unistd.h>

namespace {

using Hash = std::array<uint8_t, 32>;
struct HashHasher {
    size_t operator()(const Hash &h) const
    {
        size_t v;
        memcpy(&v, h.data() + 8, sizeof(v));  // content hashes are uniform already
        return v;
    }
};

Hash to_hash(const uint8_t *p)
{
    Hash h;
    memcpy(h.data(), p, 32);
    return h;
}

// ---------------------------------------------------------------- one node's local store
// mirrors BlockManagerLocked::write_block_inner / find_block / read_block_from /
// move_block_to_corrupted (src/block/manager.rs:720-819) with a map instead of a directory tree
struct StoredShard {
    int index = -1;          // which of the k+m shards (would be part of the file name / header, row f3)
    uint32_t block_len = 0;  // unpadded length of the whole block
    std::vector<uint8_t> bytes;
    Hash sum{};  // blake2sum of `bytes` (row f2)
};

// On-disk shard file (row f3; mirrors the tmp-file -> rename -> .corrupted life cycle of
// BlockManagerLocked::write_block_inner / move_block_to_corrupted, src/block/manager.rs:720-819,
// and the directory scheme data_dir/<h[0]>/<h[1]>/<hex(h)> of src/block/layout.rs:286-291):
//   <data_dir>/node<N>/<hh>/<hh>/<64 hex>.shard        64-byte header + shard bytes
struct ShardFileHeader {
    char magic[4];  // "GEC1"
    uint8_t k, m, index, reserved;
    uint32_t block_len;
    uint32_t shard_len;
    uint8_t sum[32];
    uint8_t pad[16];
};
static_assert(sizeof(ShardFileHeader) == 64, "shard file header is 64 bytes");

std::string hex_of(const uint8_t *p, size_t n)
{
    static const char *d = "0123456789abcdef";
    std::string o;
    for (size_t i = 0; i < n; i++) {
        o.push_back(d[p[i] >> 4]);
        o.push_back(d[p[i] & 15]);
    }
    return o;
}

struct Node {
    std::mutex mu;  // stands in for the 256 hash-sharded mutexes (manager.rs:114,679-689)
    bool up = true;
    std::string dir;  // empty: in-memory store
    int k = 0, m = 0;
    std::unordered_map<Hash, StoredShard, HashHasher> shards;     // in-memory mode
    std::unordered_map<Hash, StoredShard, HashHasher> corrupted;  // the ".corrupted" quarantine
    std::deque<Hash> resync_queue;                                // block_local_resync_queue (resync.rs:90)
    std::unordered_set<Hash, HashHasher> queued;

    std::string path_of(const Hash &h, const char *ext) const
    {
        return dir + "/" + hex_of(h.data(), 1) + "/" + hex_of(h.data() + 1, 1) + "/" + hex_of(h.data(), 32) + ext;
    }
    // all four are called with `mu` held
    bool store_put(const Hash &h, StoredShard &&s)
    {
        if (dir.empty()) {
            shards[h] = std::move(s);
            return true;
        }
        std::error_code ec;
        std::filesystem::create_directories(dir + "/" + hex_of(h.data(), 1) + "/" + hex_of(h.data() + 1, 1), ec);
        const std::string fin = path_of(h, ".shard"), tmp = fin + ".tmp" + std::to_string((unsigned long)::getpid());
        ShardFileHeader hd;
        memset(&hd, 0, sizeof(hd));
        memcpy(hd.magic, "GEC1", 4);
        hd.k = (uint8_t)k;
        hd.m = (uint8_t)m;
        hd.index = (uint8_t)s.index;
        hd.block_len = s.block_len;
        hd.shard_len = (uint32_t)s.bytes.size();
        memcpy(hd.sum, s.sum.data(), 32);
        FILE *f = fopen(tmp.c_str(), "wb");
        if (!f) return false;
        bool ok = fwrite(&hd, sizeof(hd), 1, f) == 1 &&
                  (s.bytes.empty() || fwrite(s.bytes.data(), s.bytes.size(), 1, f) == 1);
        ok = (fclose(f) == 0) && ok;
        if (ok) ok = ::rename(tmp.c_str(), fin.c_str()) == 0;  // atomic publish (manager.rs:790-795)
        if (!ok) ::remove(tmp.c_str());
        return ok;
    }
    bool store_get(const Hash &h, StoredShard &out) const
    {
        if (dir.empty()) {
            auto it = shards.find(h);
            if (it == shards.end()) return false;
            out = it->second;
            return true;
        }
        FILE *f = fopen(path_of(h, ".shard").c_str(), "rb");  // find_block (manager.rs:627-662)
        if (!f) return false;
        ShardFileHeader hd;
        bool ok = fread(&hd, sizeof(hd), 1, f) == 1 && memcmp(hd.magic, "GEC1", 4) == 0;
        if (ok) {
            out.index = hd.index;
            out.block_len = hd.block_len;
            memcpy(out.sum.data(), hd.sum, 32);
            out.bytes.resize(hd.shard_len);
            ok = hd.shard_len == 0 || fread(out.bytes.data(), hd.shard_len, 1, f) == 1;
        }
        fclose(f);
        return ok;
    }
    bool store_has(const Hash &h) const
    {
        if (dir.empty()) return shards.count(h) != 0;
        std::error_code ec;
        return std::filesystem::exists(path_of(h, ".shard"), ec);
    }
    bool store_erase(const Hash &h)
    {
        if (dir.empty()) return shards.erase(h) != 0;
        return ::remove(path_of(h, ".shard").c_str()) == 0;
    }
    void store_quarantine(const Hash &h)  // move_block_to_corrupted (manager.rs:807-819)
    {
        if (dir.empty()) {
            auto it = shards.find(h);
            if (it != shards.end()) {
                corrupted[h] = std::move(it->second);
                shards.erase(it);
            }
            return;
        }
        ::rename(path_of(h, ".shard").c_str(), path_of(h, ".corrupted").c_str());
    }
    void store_list(std::vector<Hash> &out) const  // BlockStoreIterator (repair.rs:634-764)
    {
        if (dir.empty()) {
            for (auto &kv : shards) out.push_back(kv.first);
            return;
        }
        std::error_code ec;
        for (auto it = std::filesystem::recursive_directory_iterator(dir, ec);
             !ec && it != std::filesystem::recursive_directory_iterator(); it.increment(ec)) {
            const std::string name = it->path().filename().string();
            if (name.size() != 64 + 6 || name.compare(64, 6, ".shard") != 0) continue;
            Hash h;
            bool ok = true;
            for (int i = 0; i < 32 && ok; i++) {
                auto v = [&](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : -1); };
                const int hi = v(name[2 * i]), lo = v(name[2 * i + 1]);
                ok = hi >= 0 && lo >= 0;
                h[i] = (uint8_t)((hi << 4) | lo);
            }
            if (ok) out.push_back(h);
        }
    }
};

// ---------------------------------------------------------------- counting semaphore (bytes)
// mirrors buffer_kb_semaphore (manager.rs:156, 380-385)
class ByteSemaphore {
public:
    explicit ByteSemaphore(uint64_t cap) : cap_(cap), avail_(cap) {}
    void acquire(uint64_t n)
    {
        if (n > cap_) n = cap_;
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return avail_ >= n; });
        avail_ -= n;
    }
    void release(uint64_t n)
    {
        if (n > cap_) n = cap_;
        {
            std::lock_guard<std::mutex> lk(mu_);
            avail_ += n;
        }
        cv_.notify_all();
    }

private:
    uint64_t cap_, avail_;
    std::mutex mu_;
    std::condition_variable cv_;
};

// ---------------------------------------------------------------- generic batcher (row f1)
template <class Item>
class Batcher {
public:
    using Run = std::function<void(int /*worker*/, std::vector<Item *> &)>;
    // `workers` dispatcher threads pull batches from one queue: while one batch is on the GPU the
    // next one is being collected / copied (each worker owns a garage_ec context and buffers)
    Batcher(size_t max_items, unsigned linger_us, int workers, Run run)
        : max_(std::max<size_t>(1, max_items)), linger_(linger_us), run_(std::move(run))
    {
        for (int w = 0; w < std::max(1, workers); w++) th_.emplace_back([this, w] { loop(w); });
    }
    ~Batcher()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    // blocks the calling thread until its item has been processed in some batch
    int submit(Item &it)
    {
        std::future<int> f = it.done.get_future();
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(&it);
        }
        cv_.notify_all();
        return f.get();
    }
    uint64_t batches() const { return batches_.load(); }
    uint64_t items() const { return items_.load(); }

private:
    void loop(int w)
    {
        std::vector<Item *> batch;
        for (;;) {
            batch.clear();
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (stop_ && q_.empty()) return;
                // linger: give concurrent callers a moment to join the batch
                const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(linger_);
                while (q_.size() < max_ && !stop_) {
                    if (cv_.wait_until(lk, deadline) == std::cv_status::timeout) break;
                }
                while (!q_.empty() && batch.size() < max_) {
                    batch.push_back(q_.front());
                    q_.pop_front();
                }
            }
            if (batch.empty()) continue;  // another worker took them
            batches_++;
            items_ += batch.size();
            run_(w, batch);  // sets every item's promise
        }
    }
    const size_t max_;
    const unsigned linger_;
    Run run_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Item *> q_;
    bool stop_ = false;
    std::atomic<uint64_t> batches_{0}, items_{0};
    std::vector<std::thread> th_;
};

// pool of pinned block-sized slots: the calling thread copies its block into a slot (in
// parallel with every other caller) so the dispatcher's H2D runs at PCIe speed from pinned
// memory -- in Garage proper the body copy of BytesBuf::take_exact would land here directly
class SlotPool {
public:
    bool init(garage_ec_ctx *ctx, size_t slots, size_t slot_bytes)
    {
        ctx_ = ctx;
        slot_bytes_ = slot_bytes;
        if (garage_ec_host_alloc(ctx, &base_, slots * slot_bytes) != GARAGE_EC_OK) return false;
        for (size_t i = 0; i < slots; i++) free_.push_back(static_cast<uint8_t *>(base_) + i * slot_bytes);
        return true;
    }
    uint8_t *acquire()
    {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !free_.empty(); });
        uint8_t *p = free_.back();
        free_.pop_back();
        return p;
    }
    void release(uint8_t *p)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            free_.push_back(p);
        }
        cv_.notify_one();
    }
    size_t slot_bytes() const { return slot_bytes_; }
    void destroy()
    {
        if (base_) garage_ec_host_free(ctx_, base_);
        base_ = nullptr;
    }

private:
    garage_ec_ctx *ctx_ = nullptr;
    void *base_ = nullptr;
    size_t slot_bytes_ = 0;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<uint8_t *> free_;
};

struct EncodeItem {
    const uint8_t *data = nullptr;
    uint32_t len = 0;
    std::vector<uint8_t> parity;  // m * shard_len bytes, row i at i * shard_len
    std::vector<Hash> sums;       // k+m
    std::promise<int> done;
};

struct ReconItem {
    uint32_t block_len = 0;
    std::vector<const uint8_t *> shard;          // k+m pointers, nullptr = absent
    std::vector<uint8_t> want;                   // k+m
    std::vector<std::vector<uint8_t>> rebuilt;   // k+m, filled for wanted absent shards
    std::promise<int> done;
};

// pinned scratch that grows on demand (owned by one dispatcher thread)
struct PinnedBuf {
    garage_ec_ctx *ctx = nullptr;
    void *p = nullptr;
    size_t cap = 0;
    uint8_t *get(size_t n)
    {
        if (n > cap) {
            if (p) garage_ec_host_free(ctx, p);
            p = nullptr;
            cap = 0;
            if (garage_ec_host_alloc(ctx, &p, n) != GARAGE_EC_OK) return nullptr;
            cap = n;
        }
        return static_cast<uint8_t *>(p);
    }
    void release()
    {
        if (p) garage_ec_host_free(ctx, p);
        p = nullptr;
        cap = 0;
    }
    ~PinnedBuf() { release(); }
};

}  // namespace

// ================================================================= the manager
struct garage_bm {
    garage_bm_config cfg{};
    int k = 0, m = 0, tot = 0;
    static constexpr int kWorkers = 3;  // dispatcher threads per batcher (overlap PCIe / GPU / CPU copies)
    garage_ec_ctx *ec = nullptr;                 // scrub + geometry helpers
    garage_ec_ctx *enc_ctx[kWorkers] = {nullptr};  // one context (= one set of staging lanes) per worker
    garage_ec_ctx *rec_ctx[kWorkers] = {nullptr};
    SlotPool slots;
    std::vector<std::unique_ptr<Node>> nodes;
    // stands in for the block_ref / rc tables (src/model/s3/block_ref_table.rs, src/block/rc.rs):
    // which blocks exist and how long they are
    std::mutex refs_mu;
    std::unordered_map<Hash, uint32_t, HashHasher> refs;
    // block reference counts (src/block/rc.rs): only blocks that went through block_incref /
    // block_decref have an entry; a block without entry is treated as needed (rc > 0)
    std::unordered_map<Hash, int64_t, HashHasher> rc;
    std::unique_ptr<ByteSemaphore> ram;
    std::unique_ptr<Batcher<EncodeItem>> enc_batcher;
    std::unique_ptr<Batcher<ReconItem>> rec_batcher;
    PinnedBuf enc_parity[kWorkers], rec_buf[kWorkers], scrub_buf;
    std::mutex scrub_mu;
    // metrics (src/block/metrics.rs)
    std::atomic<uint64_t> bytes_written{0}, bytes_read{0}, corruption_counter{0}, resync_counter{0},
        resync_error_counter{0}, resync_recv_counter{0}, delete_counter{0}, put_calls{0}, reconstruct_calls{0},
        scrub_checked{0}, scrub_corrupt{0}, enc_gpu_us{0}, rec_gpu_us{0};

    // rpc/layout/version.rs:117-137: top 8 bits of the hash -> partition -> k+m distinct nodes
    void storage_nodes_of(const Hash &h, int *out) const
    {
        const int n = (int)nodes.size();
        const int p = h[0] % n;
        for (int i = 0; i < tot; i++) out[i] = (p + i) % n;
    }

    void put_to_resync(int node, const Hash &h)  // manager.rs:592-605 / resync.rs:put_to_resync
    {
        Node &nd = *nodes[node];
        std::lock_guard<std::mutex> lk(nd.mu);
        if (nd.queued.insert(h).second) nd.resync_queue.push_back(h);
    }

    // manager.rs:517-530 write_block
    void write_shard(int node, const Hash &h, int index, uint32_t block_len, const uint8_t *bytes, size_t n,
                     const Hash &sum)
    {
        StoredShard s;
        s.index = index;
        s.block_len = block_len;
        s.bytes.assign(bytes, bytes + n);
        s.sum = sum;
        Node &nd = *nodes[node];
        std::lock_guard<std::mutex> lk(nd.mu);
        nd.store_put(h, std::move(s));
        bytes_written += n;
    }

    // manager.rs:554-609 read_block + read_block_from: returns a COPY of the verified shard, or false.
    // A checksum mismatch quarantines the shard and queues a resync, like the reference.
    bool read_shard(int node, const Hash &h, StoredShard &out)
    {
        Node &nd = *nodes[node];
        {
            std::lock_guard<std::mutex> lk(nd.mu);
            if (!nd.up) return false;
            if (!nd.store_get(h, out)) return false;
        }
        bytes_read += out.bytes.size();
        Hash got;
        garage_ec_blake2sum(out.bytes.data(), out.bytes.size(), got.data());
        if (got == out.sum) return true;
        corruption_counter++;
        {
            std::lock_guard<std::mutex> lk(nd.mu);
            nd.store_quarantine(h);
        }
        put_to_resync(node, h);
        return false;
    }

    // ---- batch runners (dispatcher threads) -------------------------------------------------
    void run_encode(int w, std::vector<EncodeItem *> &b)
    {
        const size_t n = b.size();
        uint32_t max_len = 0;
        std::vector<const uint8_t *> ptrs(n);
        std::vector<uint32_t> lens(n);
        for (size_t i = 0; i < n; i++) {
            ptrs[i] = b[i]->data;
            lens[i] = b[i]->len;
            max_len = std::max(max_len, b[i]->len);
        }
        const size_t stride = garage_ec_stride_for(garage_ec_shard_len(max_len, k));
        uint8_t *par = enc_parity[w].get(n * m * stride + n * tot * 32);
        int rc = par ? GARAGE_EC_OK : GARAGE_EC_E_NOMEM;
        uint8_t *sums = par ? par + n * m * stride : nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        if (rc == GARAGE_EC_OK)
            rc = garage_ec_encode_blocks_with_sums(enc_ctx[w], ptrs.data(), lens.data(), n, par, sums, stride);
        enc_gpu_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        for (size_t i = 0; i < n; i++) {
            if (rc == GARAGE_EC_OK) {
                const size_t L = garage_ec_shard_len(lens[i], k);
                b[i]->parity.resize((size_t)m * L);
                for (int r = 0; r < m; r++) memcpy(b[i]->parity.data() + r * L, par + (i * m + r) * stride, L);
                b[i]->sums.resize(tot);
                for (int s = 0; s < tot; s++) memcpy(b[i]->sums[s].data(), sums + (i * tot + s) * 32, 32);
            }
            b[i]->done.set_value(rc);
        }
    }

    void run_reconstruct(int w, std::vector<ReconItem *> &b)
    {
        const size_t n = b.size();
        uint32_t max_len = 0;
        for (auto *it : b) max_len = std::max(max_len, it->block_len);
        const size_t stride = garage_ec_stride_for(garage_ec_shard_len(max_len, k));
        const size_t data_b = n * tot * stride;
        uint8_t *buf = rec_buf[w].get(data_b + 2 * n * tot + n * 8);
        if (!buf) {
            for (auto *it : b) it->done.set_value(GARAGE_EC_E_NOMEM);
            return;
        }
        uint8_t *present = buf + data_b, *want = present + n * tot;
        std::vector<uint32_t> lens(n);
        std::vector<int32_t> status(n, 0);
        for (size_t i = 0; i < n; i++) {
            const size_t L = garage_ec_shard_len(b[i]->block_len, k);
            lens[i] = (uint32_t)L;
            for (int s = 0; s < tot; s++) {
                present[i * tot + s] = b[i]->shard[s] ? 1 : 0;
                want[i * tot + s] = b[i]->want[s];
                if (b[i]->shard[s]) memcpy(buf + (i * tot + s) * stride, b[i]->shard[s], L);
            }
        }
        const auto t0 = std::chrono::steady_clock::now();
        int rc = garage_ec_reconstruct(rec_ctx[w], buf, present, want, status.data(), lens.data(), stride, n,
                                       GARAGE_EC_MEM_HOST, nullptr);
        rec_gpu_us += (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        for (size_t i = 0; i < n; i++) {
            int r = rc;
            if (rc == GARAGE_EC_OK || rc == GARAGE_EC_E_UNRECOVERABLE) {
                r = status[i] ? GARAGE_BM_E_MISSING_BLOCK : GARAGE_BM_OK;
                if (r == GARAGE_BM_OK) {
                    b[i]->rebuilt.assign(tot, {});
                    for (int s = 0; s < tot; s++)
                        if (!b[i]->shard[s] && b[i]->want[s])
                            b[i]->rebuilt[s].assign(buf + (i * tot + s) * stride, buf + (i * tot + s) * stride + lens[i]);
                }
            }
            b[i]->done.set_value(r);
        }
    }

    // gather the valid shards of `h` from every up node except `skip_node`
    int gather(const Hash &h, int skip_node, std::vector<StoredShard> &got, std::vector<uint8_t> &have,
               uint32_t &block_len)
    {
        int who[64];
        storage_nodes_of(h, who);
        got.assign(tot, {});
        have.assign(tot, 0);
        int count = 0;
        auto try_shard = [&](int i) {
            if (who[i] == skip_node || have[i]) return;
            StoredShard s;
            if (read_shard(who[i], h, s) && s.index == i) {
                block_len = s.block_len;
                got[i] = std::move(s);
                have[i] = 1;
                count++;
            }
        };
        // data shards first (a complete set needs no GPU), then only as many parity shards as
        // it takes to reach k -- the reference likewise stops at the first good copy
        // (manager.rs:292-334); RpcHelper::try_call_many with quorum k is the real-cluster form
        for (int i = 0; i < k; i++) try_shard(i);
        for (int i = k; i < tot && count < k; i++) try_shard(i);
        return count;
    }

    int rpc_put_block(const Hash &h, const uint8_t *data, size_t len)
    {
        if (len == 0 || len > 0xffffffffull) return GARAGE_BM_E_MESSAGE;
        int who[64];
        storage_nodes_of(h, who);  // manager.rs:373
        // DataBlock::from_buffer (manager.rs:376): compression is 'none' in this mirror
        const uint64_t permits = (uint64_t)len * tot / k;  // manager.rs:380-385, x (k+m)/k for the parity
        ram->acquire(permits);
        put_calls++;
        // land the block in pinned memory (caller's thread, so copies of concurrent PUTs overlap)
        uint8_t *slot = len <= slots.slot_bytes() ? slots.acquire() : nullptr;
        if (slot) memcpy(slot, data, len);
        EncodeItem it;
        it.data = slot ? slot : data;
        it.len = (uint32_t)len;
        int rc = enc_batcher->submit(it);
        if (slot) slots.release(slot);
        if (rc != GARAGE_EC_OK) {
            ram->release(permits);
            return rc;
        }
        const size_t L = garage_ec_shard_len((uint32_t)len, k);
        std::vector<uint8_t> pad(L);
        int stored = 0;
        for (int i = 0; i < tot; i++) {  // try_write_many_sets (manager.rs:395-405): all at once
            Node &nd = *nodes[who[i]];
            {
                std::lock_guard<std::mutex> lk(nd.mu);
                if (!nd.up) continue;
            }
            const uint8_t *src;
            if (i < k) {
                const size_t off = (size_t)i * L;
                const size_t have = off < len ? std::min(L, len - off) : 0;
                if (have == L) {
                    src = data + off;
                } else {  // zero padded tail shard (put.rs:611-615 short last block)
                    std::fill(pad.begin(), pad.end(), 0);
                    if (have) memcpy(pad.data(), data + off, have);
                    src = pad.data();
                }
            } else {
                src = it.parity.data() + (size_t)(i - k) * L;
            }
            write_shard(who[i], h, i, (uint32_t)len, src, L, it.sums[i]);
            stored++;
        }
        ram->release(permits);
        {
            std::lock_guard<std::mutex> lk(refs_mu);
            refs[h] = (uint32_t)len;
        }
        const int quorum = std::min(tot, k + 1);
        return stored >= quorum ? GARAGE_BM_OK : GARAGE_BM_E_QUORUM;
    }

    int rpc_get_block(const Hash &h, std::vector<uint8_t> &out)
    {
        std::vector<StoredShard> got;
        std::vector<uint8_t> have;
        uint32_t block_len = 0;
        const int count = gather(h, -1, got, have, block_len);
        if (count < k) return GARAGE_BM_E_MISSING_BLOCK;  // manager.rs:336-338
        const size_t L = garage_ec_shard_len(block_len, k);
        bool all_data = true;
        for (int j = 0; j < k; j++) all_data &= have[j] != 0;
        std::vector<std::vector<uint8_t>> rebuilt;
        if (!all_data) {
            reconstruct_calls++;
            ReconItem it;
            it.block_len = block_len;
            it.shard.assign(tot, nullptr);
            it.want.assign(tot, 0);
            for (int i = 0; i < tot; i++) {
                if (have[i]) it.shard[i] = got[i].bytes.data();
                else if (i < k) it.want[i] = 1;
            }
            int rc = rec_batcher->submit(it);
            if (rc != GARAGE_BM_OK) return rc;
            rebuilt = std::move(it.rebuilt);
        }
        out.resize(block_len);
        for (int j = 0; j < k; j++) {
            const size_t off = (size_t)j * L;
            if (off >= block_len) break;
            const size_t n = std::min(L, (size_t)block_len - off);
            memcpy(out.data() + off, have[j] ? got[j].bytes.data() : rebuilt[j].data(), n);
        }
        return GARAGE_BM_OK;
    }

    int resync_block(int node, const Hash &h)
    {
        int who[64];
        storage_nodes_of(h, who);
        int idx = -1;
        for (int i = 0; i < tot; i++)
            if (who[i] == node) idx = i;
        if (idx < 0) return GARAGE_BM_OK;  // not a storage node for it any more (resync.rs:466-477)
        bool known, deletable;
        {
            std::lock_guard<std::mutex> lk(refs_mu);
            known = refs.count(h) != 0;
            auto it = rc.find(h);
            deletable = it != rc.end() && it->second <= 0;  // rc.is_deletable() (rc.rs)
        }
        {
            Node &nd = *nodes[node];
            std::lock_guard<std::mutex> lk(nd.mu);
            if (!nd.up) return GARAGE_BM_E_MESSAGE;
            const bool exists = nd.store_has(h);
            if (exists && deletable) {  // offload branch, resync.rs:369-458: nobody needs it -> delete_if_unneeded
                if (nd.store_erase(h)) delete_counter++;
                return GARAGE_BM_OK;
            }
            if (exists) return GARAGE_BM_OK;
        }
        if (!known || deletable) return GARAGE_BM_OK;  // rc == 0: nothing to fetch
        std::vector<StoredShard> got;
        std::vector<uint8_t> have;
        uint32_t block_len = 0;
        const int count = gather(h, node, got, have, block_len);
        if (count < k) {
            resync_error_counter++;
            return GARAGE_BM_E_MISSING_BLOCK;  // resync.rs:488-494
        }
        reconstruct_calls++;
        ReconItem it;
        it.block_len = block_len;
        it.shard.assign(tot, nullptr);
        it.want.assign(tot, 0);
        for (int i = 0; i < tot; i++)
            if (have[i]) it.shard[i] = got[i].bytes.data();
        it.want[idx] = 1;
        int rc = rec_batcher->submit(it);
        if (rc != GARAGE_BM_OK) {
            resync_error_counter++;
            return rc;
        }
        resync_recv_counter++;
        Hash sum;
        garage_ec_blake2sum(it.rebuilt[idx].data(), it.rebuilt[idx].size(), sum.data());
        write_shard(node, h, idx, block_len, it.rebuilt[idx].data(), it.rebuilt[idx].size(), sum);  // resync.rs:499
        resync_counter++;
        return GARAGE_BM_OK;
    }

    // BlockManager::block_incref / block_decref (manager.rs:452-500), driven in Garage by
    // BlockRefTable::updated (src/model/s3/block_ref_table.rs:69-85): a 0 -> 1 transition queues a
    // resync on every storage node of the block (safety check that the shard really arrives), a
    // drop to 0 queues one too (so that the shard gets deleted).
    void block_incref(const Hash &h)
    {
        bool first;
        {
            std::lock_guard<std::mutex> lk(refs_mu);
            int64_t &c = rc[h];
            first = c <= 0;
            c = first ? 1 : c + 1;
        }
        if (first) queue_on_storage_nodes(h);
    }
    void block_decref(const Hash &h)
    {
        bool zero = false;
        {
            std::lock_guard<std::mutex> lk(refs_mu);
            auto it = rc.find(h);
            if (it == rc.end()) it = rc.emplace(h, 1).first;  // an un-counted block counts as referenced once
            if (it->second > 0 && --it->second == 0) {
                zero = true;
                refs.erase(h);
            }
        }
        if (zero) queue_on_storage_nodes(h);
    }
    void queue_on_storage_nodes(const Hash &h)
    {
        int who[64];
        storage_nodes_of(h, who);
        for (int i = 0; i < tot; i++) put_to_resync(who[i], h);
    }
    int64_t get_block_rc(const Hash &h)
    {
        std::lock_guard<std::mutex> lk(refs_mu);
        auto it = rc.find(h);
        return it == rc.end() ? -1 : it->second;
    }

    int resync_all(int node, int workers, uint64_t *resynced)
    {
        Node &nd = *nodes[node];
        std::vector<Hash> todo;
        {
            std::lock_guard<std::mutex> lk(nd.mu);
            todo.assign(nd.resync_queue.begin(), nd.resync_queue.end());
            nd.resync_queue.clear();
            nd.queued.clear();
        }
        std::atomic<size_t> next{0};
        std::atomic<uint64_t> ok{0};
        std::mutex failed_mu;
        std::vector<Hash> failed;
        workers = std::max(1, std::min(workers, 64));
        std::vector<std::thread> th;
        for (int w = 0; w < workers; w++)
            th.emplace_back([&] {
                for (;;) {
                    const size_t i = next++;
                    if (i >= todo.size()) return;
                    if (resync_block(node, todo[i]) == GARAGE_BM_OK) {
                        ok++;
                    } else {
                        std::lock_guard<std::mutex> lk(failed_mu);
                        failed.push_back(todo[i]);
                    }
                }
            });
        for (auto &t : th) t.join();
        for (const Hash &h : failed) put_to_resync(node, h);  // stays queued (backoff is the caller's)
        if (resynced) *resynced = ok.load();
        return (int)failed.size();
    }

    int repair_enqueue_missing(int node, uint64_t *enq)
    {
        std::vector<Hash> all;
        {
            std::lock_guard<std::mutex> lk(refs_mu);
            for (auto &kv : refs) all.push_back(kv.first);
        }
        uint64_t c = 0;
        for (const Hash &h : all) {
            int who[64];
            storage_nodes_of(h, who);
            bool mine = false;
            for (int i = 0; i < tot; i++) mine |= who[i] == node;
            if (!mine) continue;
            bool has;
            {
                std::lock_guard<std::mutex> lk(nodes[node]->mu);
                has = nodes[node]->store_has(h);
            }
            if (!has) {
                put_to_resync(node, h);
                c++;
            }
        }
        if (enq) *enq = c;
        return GARAGE_BM_OK;
    }

    // ScrubWorker sweep of one node (repair.rs:438-490) with the GPU doing the hashing
    // `cursor`/`max_shards`: the scrub iterator checkpoint of the reference (ScrubWorker persists
    // its BlockStoreIterator position every 60 s, repair.rs:186-193,460-464, so a sweep survives a
    // restart): shards are visited in hash order, starting after *cursor (NULL = from the start),
    // at most max_shards of them (0 = all); *cursor_out receives the last hash visited and
    // *finished whether the end of the store was reached.
    int scrub(int node, uint64_t *checked, uint64_t *corrupt, const Hash *cursor = nullptr, size_t max_shards = 0,
              Hash *cursor_out = nullptr, int *finished = nullptr)
    {
        std::lock_guard<std::mutex> sl(scrub_mu);
        Node &nd = *nodes[node];
        std::vector<Hash> hashes;
        {
            std::lock_guard<std::mutex> lk(nd.mu);
            if (!nd.up) return GARAGE_BM_E_MESSAGE;
            nd.store_list(hashes);
        }
        std::sort(hashes.begin(), hashes.end());
        if (cursor) hashes.erase(hashes.begin(), std::upper_bound(hashes.begin(), hashes.end(), *cursor));
        bool done = true;
        if (max_shards && hashes.size() > max_shards) {
            hashes.resize(max_shards);
            done = false;
        }
        if (finished) *finished = done ? 1 : 0;
        if (cursor_out) {
            if (!hashes.empty()) *cursor_out = hashes.back();
            else if (cursor) *cursor_out = *cursor;
            else cursor_out->fill(0);
        }
        uint64_t nchecked = 0, nbad = 0;
        const size_t chunk = std::max<size_t>(1, (size_t)cfg.batch_max_blocks * tot);
        for (size_t c0 = 0; c0 < hashes.size(); c0 += chunk) {
            const size_t n = std::min(chunk, hashes.size() - c0);
            std::vector<StoredShard> snap(n);
            std::vector<uint8_t> ok(n, 0);
            uint32_t max_len = 0;
            {
                std::lock_guard<std::mutex> lk(nd.mu);
                for (size_t i = 0; i < n; i++) {
                    if (!nd.store_get(hashes[c0 + i], snap[i])) continue;
                    ok[i] = 1;
                    max_len = std::max<uint32_t>(max_len, (uint32_t)snap[i].bytes.size());
                }
            }
            const size_t stride = garage_ec_stride_for(max_len);
            uint8_t *buf = scrub_buf.get(n * stride + n * 32 + n + n * 4);
            if (!buf) return GARAGE_EC_E_NOMEM;
            uint8_t *expect = buf + n * stride, *bad = expect + n * 32;
            std::vector<uint32_t> lens(n, 0);
            for (size_t i = 0; i < n; i++) {
                if (!ok[i]) {
                    memset(expect + i * 32, 0, 32);
                    continue;
                }
                lens[i] = (uint32_t)snap[i].bytes.size();
                memcpy(buf + i * stride, snap[i].bytes.data(), lens[i]);
                memcpy(expect + i * 32, snap[i].sum.data(), 32);
                bytes_read += lens[i];
            }
            int rc = garage_ec_check_sums(ec, buf, expect, lens.data(), stride, n, 1, bad, GARAGE_EC_MEM_HOST, nullptr);
            if (rc != GARAGE_EC_OK) return rc;
            for (size_t i = 0; i < n; i++) {
                if (!ok[i]) continue;
                nchecked++;
                if (!bad[i]) continue;
                nbad++;
                corruption_counter++;
                {
                    std::lock_guard<std::mutex> lk(nd.mu);
                    nd.store_quarantine(hashes[c0 + i]);
                }
                put_to_resync(node, hashes[c0 + i]);
            }
        }
        scrub_checked += nchecked;
        scrub_corrupt += nbad;
        if (checked) *checked = nchecked;
        if (corrupt) *corrupt = nbad;
        return GARAGE_BM_OK;
    }
};

// ================================================================= C API
extern "C" {

void garage_bm_default_config(garage_bm_config *c)
{
    if (!c) return;
    c->data_shards = 10;
    c->parity_shards = 4;
    c->cuda_device = 0;
    c->n_nodes = 14;
    c->block_size = 1u << 20;               // util/config.rs:273-275
    c->block_ram_buffer_max = 256ull << 20;  // util/config.rs:276-278
    c->batch_max_blocks = 64;
    c->batch_linger_us = 200;
    c->data_dir = nullptr;
}

int garage_bm_create(garage_bm **out, const garage_bm_config *cfg)
{
    if (!out || !cfg) return GARAGE_EC_E_INVALID;
    *out = nullptr;
    const int k = cfg->data_shards, m = cfg->parity_shards;
    if (k < 1 || m < 1 || k + m > 40 || cfg->n_nodes < k + m || cfg->n_nodes > 256) return GARAGE_EC_E_INVALID;
    std::unique_ptr<garage_bm> bm(new garage_bm());
    bm->cfg = *cfg;
    bm->k = k;
    bm->m = m;
    bm->tot = k + m;
    int rc = garage_ec_create(&bm->ec, cfg->cuda_device, k, m, GARAGE_EC_VANDERMONDE);
    if (rc != GARAGE_EC_OK) return rc;  // no GPU => no block manager: there is no CPU fallback
    for (int i = 0; i < cfg->n_nodes; i++) {
        bm->nodes.emplace_back(new Node());
        Node &nd = *bm->nodes.back();
        nd.k = k;
        nd.m = m;
        if (cfg->data_dir && cfg->data_dir[0]) {
            nd.dir = std::string(cfg->data_dir) + "/node" + std::to_string(i);
            std::error_code ec;
            std::filesystem::create_directories(nd.dir, ec);
            // restart: what is on disk is what exists (the block_ref table would say the same)
            std::vector<Hash> have;
            nd.store_list(have);
            for (const Hash &h : have) {
                StoredShard sh;
                if (nd.store_get(h, sh)) bm->refs[h] = sh.block_len;
            }
        }
    }
    bm->ram.reset(new ByteSemaphore(cfg->block_ram_buffer_max ? cfg->block_ram_buffer_max : (256ull << 20)));
    bm->scrub_buf.ctx = bm->ec;
    for (int w = 0; w < garage_bm::kWorkers; w++) {
        rc = garage_ec_create(&bm->enc_ctx[w], cfg->cuda_device, k, m, GARAGE_EC_VANDERMONDE);
        if (rc == GARAGE_EC_OK) rc = garage_ec_create(&bm->rec_ctx[w], cfg->cuda_device, k, m, GARAGE_EC_VANDERMONDE);
        if (rc != GARAGE_EC_OK) {
            garage_bm_destroy(bm.release());
            return rc;
        }
        bm->enc_parity[w].ctx = bm->enc_ctx[w];
        bm->rec_buf[w].ctx = bm->rec_ctx[w];
    }
    const size_t nslots = (size_t)std::max<uint32_t>(cfg->batch_max_blocks, 1) * (garage_bm::kWorkers + 1);
    if (!bm->slots.init(bm->ec, nslots, cfg->block_size ? cfg->block_size : (1u << 20))) {
        garage_bm_destroy(bm.release());
        return GARAGE_EC_E_NOMEM;
    }
    garage_bm *raw = bm.get();
    bm->enc_batcher.reset(new Batcher<EncodeItem>(cfg->batch_max_blocks, cfg->batch_linger_us, garage_bm::kWorkers,
                                                  [raw](int w, std::vector<EncodeItem *> &b) { raw->run_encode(w, b); }));
    bm->rec_batcher.reset(new Batcher<ReconItem>(cfg->batch_max_blocks, cfg->batch_linger_us, garage_bm::kWorkers,
                                                 [raw](int w, std::vector<ReconItem *> &b) { raw->run_reconstruct(w, b); }));
    *out = bm.release();
    return GARAGE_BM_OK;
}

void garage_bm_destroy(garage_bm *bm)
{
    if (!bm) return;
    bm->enc_batcher.reset();
    bm->rec_batcher.reset();
    garage_ec_ctx *ec = bm->ec;
    garage_ec_ctx *ctxs[2 * garage_bm::kWorkers];
    for (int w = 0; w < garage_bm::kWorkers; w++) {
        bm->enc_parity[w].release();  // pinned buffers go before their context
        bm->rec_buf[w].release();
        ctxs[2 * w] = bm->enc_ctx[w];
        ctxs[2 * w + 1] = bm->rec_ctx[w];
    }
    bm->scrub_buf.release();
    bm->slots.destroy();
    delete bm;
    for (garage_ec_ctx *c : ctxs)
        if (c) garage_ec_destroy(c);
    garage_ec_destroy(ec);
}

void garage_bm_blake2sum(const uint8_t *data, size_t len, uint8_t hash_out[32]) { garage_ec_blake2sum(data, len, hash_out); }

int garage_bm_rpc_put_block(garage_bm *bm, const uint8_t hash[32], const uint8_t *data, size_t len)
{
    if (!bm || !hash || !data) return GARAGE_EC_E_INVALID;
    return bm->rpc_put_block(to_hash(hash), data, len);
}

int garage_bm_rpc_get_block(garage_bm *bm, const uint8_t hash[32], uint8_t *out, size_t cap, size_t *out_len)
{
    if (!bm || !hash || !out_len) return GARAGE_EC_E_INVALID;
    std::vector<uint8_t> v;
    int rc = bm->rpc_get_block(to_hash(hash), v);
    if (rc != GARAGE_BM_OK) return rc;
    *out_len = v.size();
    if (v.size() > cap || !out) return GARAGE_EC_E_INVALID;
    memcpy(out, v.data(), v.size());
    return GARAGE_BM_OK;
}

int garage_bm_resync_block(garage_bm *bm, int node, const uint8_t hash[32])
{
    if (!bm || !hash || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    return bm->resync_block(node, to_hash(hash));
}

int garage_bm_block_incref(garage_bm *bm, const uint8_t hash[32])
{
    if (!bm || !hash) return GARAGE_EC_E_INVALID;
    bm->block_incref(to_hash(hash));
    return GARAGE_BM_OK;
}

int garage_bm_block_decref(garage_bm *bm, const uint8_t hash[32])
{
    if (!bm || !hash) return GARAGE_EC_E_INVALID;
    bm->block_decref(to_hash(hash));
    return GARAGE_BM_OK;
}

long long garage_bm_get_block_rc(garage_bm *bm, const uint8_t hash[32])
{
    if (!bm || !hash) return -1;
    return (long long)bm->get_block_rc(to_hash(hash));
}

int garage_bm_resync_all(garage_bm *bm, int node, int workers, uint64_t *resynced)
{
    if (!bm || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    return bm->resync_all(node, workers, resynced);
}

int garage_bm_repair_enqueue_missing(garage_bm *bm, int node, uint64_t *enqueued)
{
    if (!bm || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    return bm->repair_enqueue_missing(node, enqueued);
}

int garage_bm_scrub(garage_bm *bm, int node, uint64_t *checked, uint64_t *corrupt)
{
    if (!bm || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    return bm->scrub(node, checked, corrupt);
}

int garage_bm_scrub_step(garage_bm *bm, int node, const uint8_t *cursor32, size_t max_shards, uint8_t cursor_out32[32],
                         int *finished, uint64_t *checked, uint64_t *corrupt)
{
    if (!bm || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    Hash cur, out;
    if (cursor32) cur = to_hash(cursor32);
    int rc = bm->scrub(node, checked, corrupt, cursor32 ? &cur : nullptr, max_shards, &out, finished);
    if (rc == GARAGE_BM_OK && cursor_out32) memcpy(cursor_out32, out.data(), 32);
    return rc;
}

int garage_bm_set_node_up(garage_bm *bm, int node, int up)
{
    if (!bm || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    std::lock_guard<std::mutex> lk(bm->nodes[node]->mu);
    bm->nodes[node]->up = up != 0;
    return GARAGE_BM_OK;
}

int garage_bm_corrupt_shard(garage_bm *bm, int node, const uint8_t hash[32], size_t byte_off)
{
    if (!bm || !hash || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    Node &nd = *bm->nodes[node];
    std::lock_guard<std::mutex> lk(nd.mu);
    StoredShard sh;
    if (!nd.store_get(to_hash(hash), sh) || sh.bytes.empty()) return GARAGE_BM_E_MISSING_BLOCK;
    sh.bytes[byte_off % sh.bytes.size()] ^= 0x01;  // the stored sum is left alone: that is the corruption
    nd.store_put(to_hash(hash), std::move(sh));
    return GARAGE_BM_OK;
}

int garage_bm_drop_shard(garage_bm *bm, int node, const uint8_t hash[32])
{
    if (!bm || !hash || node < 0 || node >= (int)bm->nodes.size()) return GARAGE_EC_E_INVALID;
    Node &nd = *bm->nodes[node];
    std::lock_guard<std::mutex> lk(nd.mu);
    if (nd.store_erase(to_hash(hash))) {
        bm->delete_counter++;
        return GARAGE_BM_OK;
    }
    return GARAGE_BM_E_MISSING_BLOCK;
}

int garage_bm_node_shard_index(garage_bm *bm, int node, const uint8_t hash[32])
{
    if (!bm || !hash || node < 0 || node >= (int)bm->nodes.size()) return -1;
    Node &nd = *bm->nodes[node];
    std::lock_guard<std::mutex> lk(nd.mu);
    StoredShard sh;
    return nd.store_get(to_hash(hash), sh) ? sh.index : -1;
}

int garage_bm_storage_nodes_of(garage_bm *bm, const uint8_t hash[32], int *nodes_out)
{
    if (!bm || !hash || !nodes_out) return GARAGE_EC_E_INVALID;
    bm->storage_nodes_of(to_hash(hash), nodes_out);
    return GARAGE_BM_OK;
}

void garage_bm_get_metrics(garage_bm *bm, garage_bm_metrics *o)
{
    if (!bm || !o) return;
    memset(o, 0, sizeof(*o));
    o->bytes_written = bm->bytes_written;
    o->bytes_read = bm->bytes_read;
    o->corruption_counter = bm->corruption_counter;
    o->resync_counter = bm->resync_counter;
    o->resync_error_counter = bm->resync_error_counter;
    o->resync_recv_counter = bm->resync_recv_counter;
    o->delete_counter = bm->delete_counter;
    o->put_calls = bm->put_calls;
    o->put_batches = bm->enc_batcher ? bm->enc_batcher->batches() : 0;
    o->reconstruct_calls = bm->reconstruct_calls;
    o->reconstruct_batches = bm->rec_batcher ? bm->rec_batcher->batches() : 0;
    o->scrub_shards_checked = bm->scrub_checked;
    o->scrub_corruptions = bm->scrub_corrupt;
    o->encode_call_us = bm->enc_gpu_us;
    o->reconstruct_call_us = bm->rec_gpu_us;
    uint64_t ql = 0;
    for (auto &n : bm->nodes) {
        std::lock_guard<std::mutex> lk(n->mu);
        ql += n->resync_queue.size();
    }
    o->resync_queue_length = ql;
}

}  // extern "C"
