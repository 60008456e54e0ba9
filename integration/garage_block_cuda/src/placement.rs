//! Row f4 in Rust: what `garage_rpc` gains so that the k+m nodes of a partition carry shard indices.
//! NOT compiled here; the executable specification is `garage_b200/csrc/placement.cpp`
//! (`include/garage_placement.h`, `tests/test_placement.py`), and every function below names its twin there.
//!
//! Reference types used as they are: `LayoutVersion`, `LayoutHelper`, `Uuid`, `Hash`
//! (src/rpc/layout/{version,helper}.rs, src/util/data.rs), `QuorumSetResultTracker` (src/rpc/rpc_helper.rs:664-760).

use std::collections::HashMap;
use std::time::Duration;

use garage_util::data::{Hash, Uuid};

/// which shard a node stores: its position in `LayoutVersion::nodes_of(hash, k + m)` (version.rs:117-137).
/// Twin: `garage_layout_nodes_of`.
pub fn shard_nodes_of(nodes_of: impl Iterator<Item = Uuid>) -> Vec<(Uuid, u8)> {
    nodes_of.enumerate().map(|(i, n)| (n, i as u8)).collect()
}

/// `ReplicationFactor::write_quorum` (replication_mode.rs:52-60) for shards: one more than the k any reader needs.
/// Twin: `garage_ec_write_quorum`.
pub fn write_quorum(k: usize, m: usize, dangerous: bool) -> usize {
    if dangerous { k } else { (k + 1).min(k + m) }
}

/// One PutShard request: `node` gets shard `index`; `sets[v]` = it counts towards the quorum of active layout
/// version v.  Twin: `garage_shard_request`.
pub struct ShardRequest { pub node: Uuid, pub index: u8, pub sets: Vec<usize> }

/// `rpc_put_block`'s `who` (manager.rs:373, `storage_sets_of`) with shards: one request per (node, index).  A node
/// that sits at different indices in the old and the new layout gets both shards; one that keeps its index (the
/// layout optimiser is asked to arrange that) gets one request that counts in both
/// sets.  Twin: `garage_layout_write_plan`.
pub fn write_plan(versions_nodes: &[Vec<Uuid>]) -> Vec<ShardRequest> {
    let mut plan: Vec<ShardRequest> = vec![];
    for (v, nodes) in versions_nodes.iter().enumerate() {
        for (i, n) in nodes.iter().enumerate() {
            match plan.iter_mut().find(|r| r.node == *n && r.index == i as u8) {
                Some(r) => r.sets.push(v),
                None => plan.push(ShardRequest { node: *n, index: i as u8, sets: vec![v] }),
            }
        }
    }
    plan
}

/// `QuorumSetResultTracker` (rpc_helper.rs:664-760) keyed by request instead of by node: same three rules
/// (success in every set >= quorum; failures + quorum > set size in any set = hopeless; success is checked first).
/// Twin: `garage_quorum_tracker_*`.
pub struct ShardQuorumTracker {
    sets_of: Vec<Vec<usize>>, answered: Vec<bool>,
    ok: Vec<usize>, err: Vec<usize>, len: Vec<usize>, quorum: usize,
}
impl ShardQuorumTracker {
    pub fn new(plan: &[ShardRequest], n_sets: usize, quorum: usize) -> Self {
        let mut len = vec![0; n_sets];
        for r in plan { for s in r.sets.iter() { len[*s] += 1; } }
        Self { sets_of: plan.iter().map(|r| r.sets.clone()).collect(), answered: vec![false; plan.len()],
               ok: vec![0; n_sets], err: vec![0; n_sets], len, quorum }
    }
    pub fn register_result(&mut self, request: usize, ok: bool) {
        if std::mem::replace(&mut self.answered[request], true) { return; }
        for s in self.sets_of[request].iter() { if ok { self.ok[*s] += 1 } else { self.err[*s] += 1 } }
    }
    pub fn all_quorums_ok(&self) -> bool { self.ok.iter().all(|c| *c >= self.quorum) }
    pub fn too_many_failures(&self) -> bool {
        self.err.iter().zip(self.len.iter()).any(|(e, l)| *e + self.quorum > *l)
    }
}

/// One place a shard can be read from.  Twin: `garage_shard_source`.
#[derive(Clone, PartialEq)]
pub struct ShardSource { pub node: Uuid, pub index: u8 }

/// `RpcHelper::request_order` (rpc_helper.rs:621-660) with "is a parity shard" in front: k data shards need no decode.
fn shard_request_order(nodes: &[Uuid], k: usize, our_node: Uuid, zone_of: &HashMap<Uuid, String>,
                       ping_of: &HashMap<Uuid, Duration>) -> Vec<ShardSource> {
    let our_zone = zone_of.get(&our_node).cloned().unwrap_or_default();
    let mut v: Vec<_> = nodes.iter().enumerate().map(|(i, n)| {
        let zone = zone_of.get(n).cloned().unwrap_or_default();
        let ping = ping_of.get(n).copied().unwrap_or_else(|| Duration::from_secs(10));
        ((i >= k, *n != our_node, zone != our_zone, ping, i), ShardSource { node: *n, index: i as u8 })
    }).collect();
    v.sort_by_key(|(key, _)| *key);
    v.into_iter().map(|(_, s)| s).collect()
}

/// `block_read_nodes_of` (rpc_helper.rs:570-619) over shard sources: active versions interleaved older to newer by
/// preference rank, ourselves first, then the historical versions.  The caller runs `try_call_many`-style
/// (rpc_helper.rs:290-411) with quorum k: start the first k sources with distinct indices; on each error start the next
/// source whose index is still missing.  Twin: `garage_layout_read_plan`.
pub fn shard_read_plan(active: &[Vec<Uuid>], historical: &[Vec<Uuid>], k: usize, our_node: Uuid,
                       zone_of: &HashMap<Uuid, String>, ping_of: &HashMap<Uuid, Duration>) -> Vec<ShardSource> {
    let ordered: Vec<Vec<ShardSource>> =
        active.iter().map(|n| shard_request_order(n, k, our_node, zone_of, ping_of)).collect();
    let mut plan: Vec<ShardSource> = vec![];
    if ordered.len() == 1 {
        plan = ordered[0].clone();
    } else {
        let rf = active.last().map(|n| n.len()).unwrap_or(0);
        for rank in 0..rf {
            for ver in ordered.iter() {
                if let Some(s) = ver.get(rank) {
                    if !plan.contains(s) {
                        if s.node == our_node { plan.insert(0, s.clone()) } else { plan.push(s.clone()) }
                    }
                }
            }
        }
    }
    for nodes in historical.iter() {
        for s in shard_request_order(nodes, k, our_node, zone_of, ping_of) {
            if !plan.contains(&s) { plan.push(s); }
        }
    }
    plan
}

/// What a layout change asks of the resync workers for one partition: shard `index` moves from `from` to `to`.  The new
/// holder first asks the old one for that very shard (a plain copy, no GPU); only if it is gone does it gather k others
/// and `EcBatcher::reconstruct(want = [index])`.  Twin: `garage_layout_transition`.
pub fn shard_moves(old_nodes: &[Uuid], new_nodes: &[Uuid]) -> Vec<(u8, Uuid, Uuid)> {
    old_nodes.iter().zip(new_nodes.iter()).enumerate()
        .filter(|(_, (a, b))| a != b).map(|(i, (a, b))| (i as u8, *a, *b)).collect()
}

/// Keep shard indices stable across a layout change: permute the new ring row so that every node that was already
/// in the partition stays at its old index (applied after `update_ring_from_flow`, version.rs:672).  Twin: the
/// `previous` argument of `garage_layout_compute`.
pub fn align_ring_row(old_row: &[u8], new_row: &mut Vec<u8>) {
    let mut out: Vec<Option<u8>> = vec![None; new_row.len()];
    let mut rest: Vec<u8> = vec![];
    for n in new_row.iter() {
        match old_row.iter().position(|o| o == n) {
            Some(i) if i < out.len() && out[i].is_none() => out[i] = Some(*n),
            _ => rest.push(*n),
        }
    }
    let mut rest = rest.into_iter();
    *new_row = out.into_iter().map(|x| x.unwrap_or_else(|| rest.next().unwrap())).collect();
}
