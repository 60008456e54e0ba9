"""Row f4: placement and quorums for erasure-coded blocks, include/garage_placement.h.
CPU only.  Every property is checked by code in this file that shares nothing with placement.cpp.

Reference behaviour the tests are modelled on: `LayoutVersion::check` (src/rpc/layout/version.rs:177-290:
distinct nodes per partition, no gateway stores data, zone redundancy, node usage within capacity),
`partition_of` (:101-104), the `QuorumSetResultTracker` rules (src/rpc/rpc_helper.rs:664-760) and the ordering
of `block_read_nodes_of` / `request_order` (:570-660)."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from garage_b200 import _build  # noqa: E402
from garage_b200 import placement as P  # noqa: E402


def h_of(partition, salt=0):
    return bytes([partition, salt & 0xff]) + bytes(30)


def independent_check(ring, zones, caps, zr, max_per_zone, partition_size):
    ring = np.asarray(ring)
    load = np.zeros(len(zones), dtype=int)
    for p in range(256):
        nodes = ring[p].tolist()
        assert len(set(nodes)) == len(nodes), ("duplicate node", p, nodes)
        assert all(caps[n] > 0 for n in nodes), ("gateway stores data", p)
        zl = {}
        for n in nodes:
            zl[zones[n]] = zl.get(zones[n], 0) + 1
            load[n] += 1
        assert len(zl) >= zr, ("zone redundancy", p, zl)
        if max_per_zone:
            assert max(zl.values()) <= max_per_zone, ("zone load", p, zl)
    for n in range(len(zones)):
        if caps[n]:
            assert load[n] <= caps[n] // partition_size, ("capacity", n, load[n])
    return load


def test_symbols_declared_bound_and_exported():
    import re

    src = open(os.path.join(ROOT, "include", "garage_placement.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(garage_(?:layout|quorum_tracker|ec)_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(P.PLACEMENT_SYMBOLS)
    P._lib()
    out = subprocess.run(["nm", "-D", "--defined-only", _build.BM_SO], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (garage_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_partition_of_is_the_top_eight_bits_of_the_hash():
    # version.rs:101-104: u16::from_be_bytes(hash[0..2]) >> (16 - PARTITION_BITS)
    rng = random.Random(1)
    for _ in range(200):
        h = bytes(rng.randrange(256) for _ in range(32))
        assert P.partition_of(h) == ((h[0] << 8 | h[1]) >> 8) == h[0]


def test_quorums():
    # reference, replicated (replication_mode.rs:44-60): rf 3 -> write quorum 2; rf 1 -> 1; rf 2 -> 2
    assert P.write_quorum(1, 2) == 2 and P.write_quorum(1, 0) == 1 and P.write_quorum(1, 1) == 2
    # shards: one more than the k a reader needs, never more than there are
    assert P.write_quorum(10, 4) == 11 and P.write_quorum(6, 3) == 7 and P.write_quorum(4, 2) == 5
    assert P.write_quorum(10, 0) == 10
    assert P.write_quorum(10, 4, P.DEGRADED) == 11 and P.write_quorum(10, 4, P.DANGEROUS) == 10
    assert P.read_quorum(10, 4) == 10 and P.read_quorum(10, 4, P.DANGEROUS) == 10
    assert P.write_quorum(0, 4) == P.E_INVALID and P.write_quorum(10, 4, 7) == P.E_INVALID


@pytest.mark.parametrize("n_nodes,n_zones,rf,max_per_zone", [(14, 1, 14, 0), (14, 4, 14, 4), (20, 5, 14, 4), (9, 3, 9, 3),
                                                              (12, 3, 6, 2), (30, 6, 14, 0), (7, 7, 6, 1)])
def test_compute_satisfies_every_constraint(n_nodes, n_zones, rf, max_per_zone):
    zones = [i % n_zones for i in range(n_nodes)]
    caps = [10**12] * n_nodes  # bytes, like the reference: the partition size is fine-grained
    lay = P.Layout.compute(zones, caps, rf, max_per_zone=max_per_zone)
    zr = min(n_zones, rf)
    load = independent_check(lay.ring(), zones, caps, zr, max_per_zone, lay.partition_size)
    rc, st = lay.check(0, max_per_zone)
    assert rc == P.OK
    assert st["storage_nodes"] == n_nodes and st["zones"] == n_zones
    assert st["min_zones_per_partition"] >= zr
    assert st["min_partitions_per_node"] == load.min() and st["max_partitions_per_node"] == load.max()
    # equal capacities: the load is even (256 * rf shards over n nodes)
    assert load.max() - load.min() <= 1
    assert load.sum() == 256 * rf
    # nodes_of = the ring row of the hash's partition; position = shard index
    for p in (0, 1, 77, 255):
        assert lay.nodes_of(h_of(p, 9)) == lay.ring()[p].tolist()


def test_capacity_proportional_and_gateways_excluded():
    zones = [0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3]
    caps = [2000, 1000, 1000, 1000] * 3 + [0]  # node 12 is a gateway
    lay = P.Layout.compute(zones, caps, 6, max_per_zone=2)
    load = independent_check(lay.ring(), zones, caps, 3, 2, lay.partition_size)
    assert load[12] == 0
    big = [load[i] for i in (0, 4, 8)]
    small = [load[i] for i in range(12) if i not in (0, 4, 8)]
    assert min(big) > max(small)  # twice the capacity buys more partitions
    # the partition size is maximal: one unit more is infeasible for the same constraints
    total = 256 * 6
    assert lay.partition_size * total <= sum(caps)
    assert all(load[n] <= caps[n] // lay.partition_size for n in range(12))


def test_infeasible_and_invalid_inputs():
    with pytest.raises(P.PlacementError) as e:
        P.Layout.compute([0] * 5, [10**9] * 5, 6)  # fewer storage nodes than shards
    assert e.value.code == P.E_INVALID
    with pytest.raises(P.PlacementError) as e:
        P.Layout.compute([0] * 8, [10**9] * 6 + [0, 0], 7)  # 6 storage nodes, 7 shards
    assert e.value.code == P.E_INFEASIBLE
    with pytest.raises(P.PlacementError) as e:
        P.Layout.compute([0, 0, 0, 1, 1, 1], [10**9] * 6, 6, zone_redundancy=3)  # only two zones
    assert e.value.code == P.E_INFEASIBLE
    with pytest.raises(P.PlacementError) as e:
        P.Layout.compute([0, 0, 0, 1, 1, 1, 2, 2], [10**9] * 8, 6, max_per_zone=1)  # 3 zones x 1 < 6
    assert e.value.code == P.E_INFEASIBLE
    with pytest.raises(P.PlacementError) as e:
        P.Layout.compute([0, 0, 0, 0, 1, 2], [10**9] * 6, 6, max_per_zone=2)  # zone 0 would need 4... only 1+1 elsewhere
    assert e.value.code == P.E_INFEASIBLE
    with pytest.raises(P.PlacementError):
        P.Layout.compute([0] * 300, [1] * 300, 3)  # compact u8 node indices


def test_zone_loss_is_survivable_exactly_when_max_per_zone_is_at_most_m():
    """RS(10,4) over 4 zones with max_per_zone = 4: losing any one zone leaves >= 10 shards of every partition;
    without the limit the optimiser is free to put 5 shards in one zone"""
    zones = [i % 4 for i in range(16)]
    caps = [10**9] * 16
    lay = P.Layout.compute(zones, caps, 14, max_per_zone=4)
    ring = lay.ring()
    for dead in range(4):
        for p in range(256):
            alive = [n for n in ring[p] if zones[n] != dead]
            assert len(alive) >= 10
    # a ring that violates it is reported by check()
    bad = ring.copy()
    z0 = [n for n in range(16) if zones[n] == 0]          # 4 nodes of zone 0
    others = [n for n in range(16) if zones[n] != 0]
    bad[5] = z0 + others[:10]
    bad[5][4] = others[10]
    lay2 = P.Layout.from_ring(zones, caps, 14, bad)
    assert lay2.check(0, 4)[0] == P.OK  # still 4 in zone 0
    bad[6] = z0 + others[:10]
    zones5 = list(zones)
    zones5[others[0]] = 0                                  # a fifth node joins zone 0
    lay3 = P.Layout.from_ring(zones5, caps, 14, bad)
    rc, st = lay3.check(0, 4)
    assert rc == P.E_ZONE_LOAD and st["max_shards_per_zone"] == 5


def test_check_reports_each_violation():
    zones = [0, 0, 1, 1, 2, 2]
    caps = [10**9] * 6
    lay = P.Layout.compute(zones, caps, 3)
    ring = lay.ring()
    dup = ring.copy()
    dup[3] = [dup[3][0], dup[3][0], dup[3][2]]
    assert P.Layout.from_ring(zones, caps, 3, dup).check()[0] == P.E_DUPLICATE
    gw = list(caps)
    gw[int(ring[0][0])] = 0
    assert P.Layout.from_ring(zones, gw, 3, ring).check()[0] == P.E_GATEWAY
    same = ring.copy()
    same[9] = [0, 1, 2]  # two zones only
    assert P.Layout.from_ring(zones, caps, 3, same).check()[0] == P.E_ZONES
    assert P.Layout.from_ring(zones, caps, 3, same).check(zone_redundancy=2)[0] == P.OK
    with pytest.raises(P.PlacementError):
        bad = ring.copy()
        bad[0][0] = 6  # no such node
        P.Layout.from_ring(zones, caps, 3, bad)


def test_layout_change_keeps_shard_indices_and_moves_little():
    zones = [i % 4 for i in range(20)]
    caps = [10**9] * 20
    v1 = P.Layout.compute(zones, caps, 14, max_per_zone=4, version=1)
    r1 = v1.ring()
    # node 7 leaves (becomes a gateway)
    caps2 = list(caps)
    caps2[7] = 0
    v2 = P.Layout.compute(zones, caps2, 14, max_per_zone=4, previous=v1, version=2)
    r2 = v2.ring()
    independent_check(r2, zones, caps2, 4, 4, v2.partition_size)
    on7 = int((r1 == 7).sum())
    moved = v1.transition_to(v2)
    assert moved == int((r1 != r2).sum())
    assert on7 <= moved <= 3 * on7, (on7, moved)
    # a node that stays in a partition keeps its shard index: it never has to swap one shard for another
    for p in range(256):
        for i in range(14):
            n = r1[p][i]
            if n in r2[p]:
                assert list(r2[p]).index(n) == i, (p, i, n)
        for idx, a, b in v1.transition_to(v2, p):
            assert r1[p][idx] == a and r2[p][idx] == b and a != b
    # recomputing without `previous` shuffles far more
    v2b = P.Layout.compute(zones, caps2, 14, max_per_zone=4, version=2)
    assert v1.transition_to(v2b) > 4 * moved
    # a node joins: only what it takes over moves
    zones3, caps3 = zones + [0], caps2 + [10**9]
    with pytest.raises(P.PlacementError):
        P.Layout.compute(zones3, caps3, 14, previous=v2)  # node numbering must match the previous version


def _tracker_model(requests, n_sets, quorum, events):
    """QuorumSetResultTracker (rpc_helper.rs:664-760) over requests instead of nodes"""
    ok, err = [0] * n_sets, [0] * n_sets
    lens = [sum(1 for _, _, m in requests if m >> s & 1) for s in range(n_sets)]
    state = P.QUORUM_PENDING
    if any(l < quorum for l in lens):
        state = P.QUORUM_FAILED
    seen = set()
    out = []
    for r, good in events:
        if r not in seen:
            seen.add(r)
            for s in range(n_sets):
                if requests[r][2] >> s & 1:
                    if good:
                        ok[s] += 1
                    else:
                        err[s] += 1
            if state == P.QUORUM_PENDING:
                if all(o >= quorum for o in ok):
                    state = P.QUORUM_OK
                elif any(e + quorum > l for e, l in zip(err, lens)):
                    state = P.QUORUM_FAILED
        out.append(state)
    return out


def test_write_plan_and_quorum_sets_across_layout_versions():
    zones = [i % 4 for i in range(20)]
    caps = [10**9] * 20
    v1 = P.Layout.compute(zones, caps, 14, max_per_zone=4, version=1)
    caps2 = list(caps)
    caps2[3] = 0
    caps2[11] = 0
    v2 = P.Layout.compute(zones, caps2, 14, max_per_zone=4, previous=v1, version=2)
    rng = random.Random(5)
    for p in (0, 13, 200):
        h = h_of(p)
        one = P.write_plan([v1], h)
        assert one == [(n, i, 1) for i, n in enumerate(v1.nodes_of(h))]
        both = P.write_plan([v1, v2], h)
        n1, n2 = v1.nodes_of(h), v2.nodes_of(h)
        expect = {}
        for v, nodes in enumerate((n1, n2)):
            for i, n in enumerate(nodes):
                expect[(n, i)] = expect.get((n, i), 0) | 1 << v
        assert {(n, i): m for n, i, m in both} == expect and len(both) == len(expect)
        # every set has exactly k+m members
        for s in range(2):
            assert sum(1 for _, _, m in both if m >> s & 1) == 14
        q = P.write_quorum(10, 4)
        for _ in range(30):
            order = list(range(len(both)))
            rng.shuffle(order)
            pfail = rng.choice([0.0, 0.1, 0.3, 0.6])
            events = [(r, rng.random() >= pfail) for r in order]
            events += [(events[0][0], not events[0][1])]  # a duplicate answer is ignored
            t = P.QuorumTracker(both, 2, q)
            got = [t.register(r, good) for r, good in events]
            assert got == _tracker_model(both, 2, q, events)
            assert t.state == got[-1]
            t.close()
    # 11 of 14 acks in the old set are not enough if the new set is short of its quorum
    h = h_of(13)
    both = P.write_plan([v1, v2], h)
    t = P.QuorumTracker(both, 2, 11)
    only_old = [j for j, (_, _, m) in enumerate(both) if m == 1]
    shared = [j for j, (_, _, m) in enumerate(both) if m == 3]
    for j in shared[:11]:
        st = t.register(j, True)
    if len(shared) >= 11:
        assert st == P.QUORUM_OK  # shared requests count in both sets
    t.close()
    t = P.QuorumTracker(both, 2, 11)
    for j in (shared + only_old)[:14]:
        st = t.register(j, True)
    new_only = [j for j, (_, _, m) in enumerate(both) if m == 2]
    if len(shared) < 11:
        assert st == P.QUORUM_PENDING  # old set complete, new set still short
        for j in new_only:
            st = t.register(j, True)
        assert st == P.QUORUM_OK
    t.close()
    with pytest.raises(P.PlacementError):
        P.QuorumTracker([(0, 0, 4)], 2, 1)  # a mask bit beyond the sets


def test_a_set_smaller_than_the_quorum_fails_immediately():
    t = P.QuorumTracker([(0, 0, 1), (1, 1, 1), (2, 2, 3)], 2, 2)  # set 1 has one member, quorum 2
    assert t.state == P.QUORUM_FAILED
    t.close()


def test_read_plan_single_version_orders_data_first_then_self_zone_ping():
    zones = [i % 4 for i in range(16)]
    caps = [10**9] * 16
    lay = P.Layout.compute(zones, caps, 14, max_per_zone=4)
    h = h_of(42)
    nodes = lay.nodes_of(h)
    k = 10
    plan = P.read_plan([lay], h, k)
    assert sorted((n, i) for n, i, _ in plan) == sorted((n, i) for i, n in enumerate(nodes))
    assert [i for _, i, _ in plan[:k]] == list(range(k))          # no locality information: data shards in index order
    assert all(i >= k for _, i, _ in plan[k:])
    our = nodes[12]                                                 # we hold a parity shard
    ping = [1000 + 10 * n for n in range(16)]
    plan = P.read_plan([lay], h, k, our_node=our, ping_us=ping)
    data, parity = plan[:k], plan[k:]
    assert all(i < k for _, i, _ in data) and parity[0][0] == our    # parity: ourselves first
    key = lambda n: (n != our, zones[n] != zones[our], ping[n])     # noqa: E731
    assert [n for n, _, _ in data] == sorted((nodes[i] for i in range(k)), key=key)
    assert [n for n, _, _ in parity] == sorted((nodes[i] for i in range(k, 14)), key=key)
    our = nodes[3]                                                  # we hold a data shard: it comes first of all
    plan = P.read_plan([lay], h, k, our_node=our, ping_us=ping)
    assert plan[0][:2] == (our, 3)


def test_read_plan_during_a_layout_change_and_with_historical_versions():
    zones = [i % 4 for i in range(20)]
    caps = [10**9] * 20
    v1 = P.Layout.compute(zones, caps, 14, max_per_zone=4, version=1)
    caps2 = list(caps)
    caps2[5] = 0
    v2 = P.Layout.compute(zones, caps2, 14, max_per_zone=4, previous=v1, version=2)
    caps3 = list(caps2)
    caps3[9] = 0
    v3 = P.Layout.compute(zones, caps3, 14, max_per_zone=4, previous=v2, version=3)
    k = 10
    for p in range(0, 256, 17):
        h = h_of(p)
        n2, n3 = v2.nodes_of(h), v3.nodes_of(h)
        plan = P.read_plan([v2, v3], h, k, old=[v1])
        srcs = [(n, i) for n, i, _ in plan]
        assert len(set(srcs)) == len(srcs)                          # no (node, index) twice
        want = {(n, i) for i, n in enumerate(n2)} | {(n, i) for i, n in enumerate(n3)} | {(n, i) for i, n in enumerate(v1.nodes_of(h))}
        assert set(srcs) == want
        # active versions first, interleaved by rank with the older one first; historical ones at the end
        vers = [v for _, _, v in plan]
        assert vers == sorted(vers, key=lambda v: v == 2) and all(v in (0, 1, 2) for v in vers)
        active = [(n, i, v) for n, i, v in plan if v < 2]
        assert active[0] == (n2[0], 0, 0)
        # the first k sources cover k distinct data indices: a healthy GET needs no decode
        first = []
        for n, i, _ in plan:
            if i not in [j for _, j in first]:
                first.append((n, i))
            if len(first) == k:
                break
        assert sorted(i for _, i in first) == list(range(k))
        # ourselves first, wherever we are
        our = n3[11]
        plan = P.read_plan([v2, v3], h, k, our_node=our, old=[v1])
        assert plan[0][0] == our
    with pytest.raises(P.PlacementError):
        P.read_plan([v2, v3], h_of(1), 15)  # k larger than the replication factor


def test_gather_simulation_any_m_failures_still_decodable():
    """walk the read plan the way the gatherer does (first k distinct indices, replacement on error): with up to m
    dead nodes every block is still assembled, and with all nodes alive no parity shard is touched"""
    zones = [i % 4 for i in range(16)]
    caps = [10**9] * 16
    lay = P.Layout.compute(zones, caps, 14, max_per_zone=4)
    rng = random.Random(3)
    k, m = 10, 4
    for p in range(0, 256, 5):
        h = h_of(p)
        plan = P.read_plan([lay], h, k, our_node=rng.randrange(16))
        for dead_count in (0, 1, 4):
            dead = set(rng.sample(range(16), dead_count))
            got = {}
            for n, i, _ in plan:
                if len(got) == k:
                    break
                if i in got or n in dead:
                    continue
                got[i] = n
            assert len(got) == k
            if dead_count == 0:
                assert sorted(got) == list(range(k))
        dead = set(rng.sample(lay.nodes_of(h), m + 1))  # one too many
        got = {i for n, i, _ in plan if n not in dead}
        assert len(got) == k - 1 + 0 or len(got) < k + m - m  # k+m-(m+1) = k-1 shards are left
        assert len(got) == k - 1


def test_the_mirror_ring_is_the_rule_it_replaced():
    """block_manager.cpp fills a garage_layout with partition p -> nodes (p mod n) + i and asks it for the nodes of a
    hash; that is node (hash[0] mod n + i) mod n for shard i, the rule the mirror used before it had a layout"""
    rng = random.Random(11)
    for n, tot in ((14, 14), (17, 14), (9, 6), (256, 36), (6, 6)):
        ring = np.array([[(p % n + i) % n for i in range(tot)] for p in range(256)], dtype=np.uint8 if n <= 256 else None)
        lay = P.Layout.from_ring([0] * n, [1] * n, tot, ring)
        assert lay.check()[0] == P.OK
        for _ in range(50):
            h = bytes(rng.randrange(256) for _ in range(32))
            assert lay.nodes_of(h) == [(h[0] % n + i) % n for i in range(tot)]


@pytest.mark.parametrize("seed", range(12))
def test_random_clusters_and_layout_changes(seed):
    """random zones / capacities / codes: every computed ring satisfies the constraints (independent checker), the load
    follows the capacities, and a follow-up layout after nodes left or joined keeps shard indices and moves at most a
    small multiple of what the departed nodes held"""
    rng = random.Random(1000 + seed)
    n = rng.randrange(8, 40)
    nz = rng.randrange(1, 7)
    rf = rng.choice([3, 6, 9, 14])
    if rf > n:
        rf = n
    zones = [rng.randrange(nz) for _ in range(n)]
    caps = [rng.choice([0, 1, 1, 2, 4]) * 10**12 for _ in range(n)]
    storage = [i for i in range(n) if caps[i]]
    real_zones = len({zones[i] for i in storage})
    mpz = rng.choice([0, 0, max(1, -(-rf // max(real_zones, 1)) + 1)])
    try:
        v1 = P.Layout.compute(zones, caps, rf, max_per_zone=mpz)
    except P.PlacementError as e:
        assert e.code == P.E_INFEASIBLE
        return  # (too few storage nodes, a zone limit that cannot be met, or capacities too skewed)
    zr = min(real_zones, rf)
    load = independent_check(v1.ring(), zones, caps, zr, mpz, v1.partition_size)
    assert load.sum() == 256 * rf
    # load follows capacity: nobody is more than one partition-size step above its share of the bytes
    for i in storage:
        assert load[i] <= caps[i] // v1.partition_size
    # change: one storage node leaves, one gateway (if any) gets capacity
    caps2 = list(caps)
    gone = rng.choice(storage)
    caps2[gone] = 0
    gw = [i for i in range(n) if caps[i] == 0]
    if gw:
        caps2[rng.choice(gw)] = 2 * 10**12
    try:
        v2 = P.Layout.compute(zones, caps2, rf, max_per_zone=mpz, previous=v1, version=2)
    except P.PlacementError as e:
        assert e.code == P.E_INFEASIBLE
        return
    real_zones2 = len({zones[i] for i in range(n) if caps2[i]})
    independent_check(v2.ring(), zones, caps2, min(real_zones2, rf), mpz, v2.partition_size)
    r1, r2 = v1.ring(), v2.ring()
    for p in range(256):
        for i in range(rf):
            if r1[p][i] in r2[p]:
                assert list(r2[p]).index(r1[p][i]) == i
    assert v1.transition_to(v2) == int((r1 != r2).sum())
