//! Batching front-end in front of `rpc_put_block` / degraded GET / the resync workers
//!.  NOT compiled here (no rustc); it is the Rust form of the
//! `Batcher` + slot pools of garage_b200/csrc/block_manager.cpp, which IS compiled, tested and
//! measured (tools/bm_bench.py).
//!
//! Shape of the problem in the reference: up to 3 blocks in flight per PUT request
//! (src/api/s3/put.rs:42), 8 resync workers (src/block/resync.rs:43), CPU-heavy work wrapped in
//! `spawn_blocking` (src/block/block.rs:86).  One block per FFI call cannot amortise a kernel launch
//! plus two PCIe transfers, so callers `submit()` and await; dispatcher threads (plain OS threads,
//! bound to the GPU's NUMA node with garage_ec_bind_thread) collect what arrived within `linger`
//! and hand the library whole batches.  Bytes are moved by the CALLERS, in parallel:
//!   * PUT: the caller copies its block into a pinned slot (where BytesBuf::take_exact,
//!     src/net/bytes_buf.rs:66-117, would land the body directly) and gets back `EncodedBlock`,
//!     whose parity shards are slices of the dispatcher's pinned output buffer (kept alive by Arc);
//!   * degraded GET / resync: the caller fills a pinned stripe slot with the shards that arrived
//!     and the batch is a list of slot pointers (garage_ec_reconstruct_stripes).
use std::sync::Arc;
use std::time::Duration;

use bytes::Bytes;
use tokio::sync::{mpsc, oneshot, Semaphore};

use garage_util::data::Hash;
use garage_util::error::Error;

use crate::{check, sys, ErasureCoder};

/// pinned host memory from garage_ec_host_alloc (NUMA-local to the GPU), freed on drop
pub struct Pinned { ec: Arc<ErasureCoder>, ptr: *mut u8, len: usize }
unsafe impl Send for Pinned {}
unsafe impl Sync for Pinned {}
impl Pinned {
    pub fn new(ec: Arc<ErasureCoder>, len: usize) -> Result<Self, Error> {
        let mut p: *mut std::ffi::c_void = std::ptr::null_mut();
        check(ec.ctx, unsafe { sys::garage_ec_host_alloc(ec.ctx, &mut p, len) })?;
        Ok(Self { ec, ptr: p as *mut u8, len })
    }
    pub fn as_slice(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
    #[allow(clippy::mut_from_ref)]
    pub unsafe fn as_mut_slice(&self) -> &mut [u8] { std::slice::from_raw_parts_mut(self.ptr, self.len) }
}
impl Drop for Pinned {
    fn drop(&mut self) { unsafe { sys::garage_ec_host_free(self.ec.ctx, self.ptr as *mut _) } }
}
impl AsRef<[u8]> for Pinned { fn as_ref(&self) -> &[u8] { self.as_slice() } }

/// what `rpc_put_block` sends out: shard i goes to node i of the (k+m)-node set
pub struct EncodedBlock {
    pub block_len: usize,
    pub shard_len: usize,
    /// k data shards (slices of the caller's block, zero padded tail materialised) then m parity shards
    pub shards: Vec<Bytes>,
    /// per-shard integrity tag of the context's sum kind (adler8 by default), k+m x 32 bytes
    pub sums: Vec<[u8; 32]>,
}

struct EncodeReq { block: Bytes, reply: oneshot::Sender<Result<EncodedBlock, Error>> }
struct ReconReq {
    hash: Hash,
    block_len: usize,
    stripe: Pinned,            // k+m shards, `stride` apart; survivors filled in by the caller
    present: Vec<u8>,
    want: Vec<u8>,
    reply: oneshot::Sender<Result<Pinned, Error>>,
}

#[derive(Clone)]
pub struct BatchConfig { pub max_blocks: usize, pub linger: Duration, pub dispatchers: usize, pub block_size: usize,
                         /// dispatchers sleep on an event instead of spinning in the driver while their batch is on the GPU
                         /// (frees one CPU per dispatcher; measured 15 % lower PUT rate on a 16-CPU host)
                         pub sleep_wait: bool }
impl Default for BatchConfig {
    fn default() -> Self { Self { max_blocks: 64, linger: Duration::from_micros(100), dispatchers: 3, block_size: 1 << 20, sleep_wait: false } }
}

/// cloneable handle held by BlockManager (next to `buffer_kb_semaphore`, src/block/manager.rs:156)
#[derive(Clone)]
pub struct EcBatcher {
    ec: Arc<ErasureCoder>,
    enc_tx: mpsc::UnboundedSender<EncodeReq>,
    rec_tx: mpsc::UnboundedSender<ReconReq>,
    /// bounds the pinned slots handed out (back-pressure like block_ram_buffer_max, manager.rs:380-385)
    slots: Arc<Semaphore>,
    pub stride: usize,
}

impl EcBatcher {
    pub fn new(ec: Arc<ErasureCoder>, cfg: BatchConfig) -> Self {
        let stride = unsafe { sys::garage_ec_stride_for(ec.shard_len(cfg.block_size) as u32) };
        ec.set_wait_blocking(cfg.sleep_wait);
        let (enc_tx, enc_rx) = mpsc::unbounded_channel::<EncodeReq>();
        let (rec_tx, rec_rx) = mpsc::unbounded_channel::<ReconReq>();
        let enc_rx = Arc::new(std::sync::Mutex::new(enc_rx));
        let rec_rx = Arc::new(std::sync::Mutex::new(rec_rx));
        for _ in 0..cfg.dispatchers {
            let (ec2, rx, c) = (ec.clone(), enc_rx.clone(), cfg.clone());
            std::thread::spawn(move || encode_dispatcher(ec2, rx, c, stride));
            let (ec2, rx, c) = (ec.clone(), rec_rx.clone(), cfg.clone());
            std::thread::spawn(move || recon_dispatcher(ec2, rx, c, stride));
        }
        let slots = Arc::new(Semaphore::new(cfg.max_blocks * (cfg.dispatchers + 1)));
        Self { ec, enc_tx, rec_tx, slots, stride }
    }

    /// ENCODE call site: BlockManager::rpc_put_block, right after DataBlock::from_buffer
    /// (src/block/manager.rs:376).  `block` should already live in pinned memory (Bytes::from_owner of
    /// a `Pinned`); a pageable Bytes works, only slower.
    pub async fn encode(&self, block: Bytes) -> Result<EncodedBlock, Error> {
        let _permit = self.slots.acquire().await.map_err(|_| Error::Message("ec batcher closed".into()))?;
        let (tx, rx) = oneshot::channel();
        self.enc_tx.send(EncodeReq { block, reply: tx }).map_err(|_| Error::Message("ec batcher closed".into()))?;
        rx.await.map_err(|_| Error::Message("ec dispatcher died".into()))?
    }

    /// DECODE call sites: rpc_get_raw_block_internal (manager.rs:276-339, want = absent data shards)
    /// and resync_block's fetch branch (resync.rs:460-500, want = this node's shard).  Returns the
    /// stripe slot with the wanted shards filled in.
    pub async fn reconstruct(&self, hash: Hash, block_len: usize, arrived: Vec<Option<Bytes>>, want: Vec<bool>)
        -> Result<Pinned, Error>
    {
        let tot = self.ec.k + self.ec.m;
        let l = self.ec.shard_len(block_len);
        let _permit = self.slots.acquire().await.map_err(|_| Error::Message("ec batcher closed".into()))?;
        let stripe = Pinned::new(self.ec.clone(), tot * self.stride)?;
        let mut present = vec![0u8; tot];
        for (i, s) in arrived.iter().enumerate() {
            if let Some(s) = s {
                // the copy into pinned memory happens on the caller's task, in parallel with all others
                // non-temporal copy: the DMA that follows reads DRAM, not this core's cache (7x on the measured hosts)
                unsafe {
                    sys::garage_ec_copy_for_dma(stripe.as_mut_slice()[i * self.stride..].as_mut_ptr() as *mut _,
                                                s.as_ptr() as *const _, l)
                };
                present[i] = 1;
            }
        }
        let want: Vec<u8> = want.iter().map(|w| *w as u8).collect();
        let (tx, rx) = oneshot::channel();
        self.rec_tx.send(ReconReq { hash, block_len, stripe, present, want, reply: tx })
            .map_err(|_| Error::Message("ec batcher closed".into()))?;
        rx.await.map_err(|_| Error::Message("ec dispatcher died".into()))?
    }
}

/// collect up to `max` requests: block for the first, then give the rest `linger` to arrive
fn collect<T>(rx: &std::sync::Mutex<mpsc::UnboundedReceiver<T>>, max: usize, linger: Duration) -> Option<Vec<T>> {
    let mut rx = rx.lock().unwrap();
    let first = rx.blocking_recv()?;
    let mut batch = vec![first];
    let deadline = std::time::Instant::now() + linger;
    while batch.len() < max {
        match rx.try_recv() {
            Ok(r) => batch.push(r),
            Err(_) if std::time::Instant::now() < deadline => std::thread::yield_now(),
            Err(_) => break,
        }
    }
    Some(batch)
}

fn encode_dispatcher(ec: Arc<ErasureCoder>, rx: Arc<std::sync::Mutex<mpsc::UnboundedReceiver<EncodeReq>>>,
                     cfg: BatchConfig, stride: usize) {
    unsafe { sys::garage_ec_bind_thread(ec.ctx) };
    let (k, m) = (ec.k, ec.m);
    while let Some(batch) = collect(&rx, cfg.max_blocks, cfg.linger) {
        let n = batch.len();
        let ptrs: Vec<*const u8> = batch.iter().map(|r| r.block.as_ptr()).collect();
        let lens: Vec<u32> = batch.iter().map(|r| r.block.len() as u32).collect();
        // one pinned output buffer per batch, shared (Arc) by the EncodedBlocks cut out of it
        let out = match Pinned::new(ec.clone(), n * m * stride + n * (k + m) * 32) {
            Ok(p) => p,
            Err(_) => { for r in batch { let _ = r.reply.send(Err(Error::Message("pinned alloc failed".into()))); } continue; }
        };
        let rc = unsafe {
            let o = out.as_mut_slice();
            let (par, sums) = o.split_at_mut(n * m * stride);
            sys::garage_ec_encode_blocks_with_sums(ec.ctx, ptrs.as_ptr(), lens.as_ptr(), n, par.as_mut_ptr(),
                                                   sums.as_mut_ptr(), stride)
        };
        let res = check(ec.ctx, rc);
        let out = Bytes::from_owner(out);
        for (s, r) in batch.into_iter().enumerate() {
            let reply = match &res {
                Err(e) => Err(Error::Message(format!("{}", e))),
                Ok(()) => {
                    let bl = r.block.len();
                    let l = ec.shard_len(bl);
                    let mut shards = Vec::with_capacity(k + m);
                    for j in 0..k {   // data shards: slices of the block; only a short tail is copied + padded
                        let (a, b) = ((j * l).min(bl), ((j + 1) * l).min(bl));
                        if b - a == l { shards.push(r.block.slice(a..b)); }
                        else { let mut v = vec![0u8; l]; v[..b - a].copy_from_slice(&r.block[a..b]); shards.push(Bytes::from(v)); }
                    }
                    for i in 0..m { let o = (s * m + i) * stride; shards.push(out.slice(o..o + l)); }
                    let so = n * m * stride + s * (k + m) * 32;
                    let sums = (0..k + m).map(|i| { let mut t = [0u8; 32]; t.copy_from_slice(&out[so + i * 32..so + i * 32 + 32]); t }).collect();
                    Ok(EncodedBlock { block_len: bl, shard_len: l, shards, sums })
                }
            };
            let _ = r.reply.send(reply);
        }
    }
}

fn recon_dispatcher(ec: Arc<ErasureCoder>, rx: Arc<std::sync::Mutex<mpsc::UnboundedReceiver<ReconReq>>>,
                    cfg: BatchConfig, stride: usize) {
    unsafe { sys::garage_ec_bind_thread(ec.ctx) };
    let tot = ec.k + ec.m;
    while let Some(batch) = collect(&rx, cfg.max_blocks, cfg.linger) {
        let n = batch.len();
        let stripes: Vec<*mut u8> = batch.iter().map(|r| r.stripe.ptr).collect();
        let lens: Vec<u32> = batch.iter().map(|r| ec.shard_len(r.block_len) as u32).collect();
        let mut present = Vec::with_capacity(n * tot);
        let mut want = Vec::with_capacity(n * tot);
        for r in &batch { present.extend_from_slice(&r.present); want.extend_from_slice(&r.want); }
        let mut status = vec![0i32; n];
        let rc = unsafe { sys::garage_ec_reconstruct_stripes(ec.ctx, stripes.as_ptr(), present.as_ptr(), want.as_ptr(),
                                                             status.as_mut_ptr(), lens.as_ptr(), stride, n) };
        for (s, r) in batch.into_iter().enumerate() {
            let reply = if status[s] != 0 { Err(Error::MissingBlock(r.hash)) }          // < k shards: resync backs off
                        else if rc != sys::GARAGE_EC_OK && rc != sys::GARAGE_EC_E_UNRECOVERABLE { check(ec.ctx, rc).map(|_| unreachable!()) }
                        else { Ok(r.stripe) };
            let _ = r.reply.send(reply);
        }
    }
}
