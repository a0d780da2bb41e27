"""N > 1 host logic on CPU: the multi-rank merge choreography of nimble_aligner_b200/multigpu.py (the one bench.py runs
over NCCL with device shards) at world_size 2 over gloo, with a host model of the per-rank tables fed by the oracle's
per-pair results.  Checks that sharded alignment + merge gives exactly the single-process counts over the union, incl.
read_keys duplicated across ranks (the later pair wins, src/align.rs:685) and callsets only one rank has seen."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
import synth
from nimble_aligner_b200.multigpu import merge_across_ranks, merge_scoped_across_ranks

GCAP = 16


def _tag(cs):
    return int.from_bytes(hashlib.blake2b(repr(cs).encode(), digest_size=8).digest(), "little") | 1


class ModelShard:
    """Host model of one rank's tables: read_key -> (global pair order, callset tag) with 'later insertable pair wins',
    callset dictionary tag -> tuple of group names (names stand in for the group ranks of the device rows)."""

    def __init__(self, names_sorted):
        self.gid = {n: i for i, n in enumerate(names_sorted)}
        self.names = names_sorted
        self.keys, self.cs = {}, {}

    def add_pair(self, key, order, insertable, callset):
        if not insertable:
            return
        tag = 0
        if callset is not None:
            tag = _tag(callset)
            self.cs[tag] = callset
        if key not in self.keys or self.keys[key][0] < order:
            self.keys[key] = (order, tag)

    def callsets_export(self):
        rows = np.zeros((len(self.cs), 4 + GCAP), dtype=np.uint32)
        for i, (tag, cs) in enumerate(sorted(self.cs.items())):
            rows[i, :4] = (i, len(cs), tag & 0xFFFFFFFF, tag >> 32)
            rows[i, 4:4 + len(cs)] = [self.gid[n] for n in cs]
        return rows

    def keys_export_partitioned(self, world):
        by = [[] for _ in range(world)]
        for key, (order, tag) in self.keys.items():
            lo, hi = key & (2 ** 64 - 1), key >> 64
            by[((lo >> 40) & 0xFFFF) % world].append((lo, hi, order, tag))
        flat = [r for b in by for r in b]
        rec = torch.from_numpy(np.array(flat, dtype=np.uint64).reshape(-1, 4).view(np.int64))
        return rec, [len(b) for b in by]

    def callsets_import(self, rows):
        for r in rows:
            tag = int(r[2]) | (int(r[3]) << 32)
            self.cs[tag] = tuple(self.names[g] for g in r[4:4 + int(r[1])])

    def recv_buffer(self, n):
        return torch.empty((max(n, 1), 4), dtype=torch.int64)

    def keys_import(self, rec):
        self.keys = {}   # after the exchange a rank only answers for the keys it owns
        for lo, hi, order, tag in rec.numpy().view(np.uint64).tolist():
            key = lo | (hi << 64)
            if key not in self.keys or self.keys[key][0] < order:
                self.keys[key] = (order, tag)

    def finalize(self):
        order = sorted(self.cs.items(), key=lambda kv: kv[1])          # same dictionary everywhere -> same dense order
        dense = {tag: i for i, (tag, _) in enumerate(order)}
        cnt = np.zeros(len(order), dtype=np.int64)
        for _, tag in self.keys.values():
            if tag:
                cnt[dense[tag]] += 1
        nz = np.flatnonzero(cnt)
        return dict(callset_off=np.arange(len(order) + 1), row_callset=nz, row_count=cnt[nz], n_unique_keys=len(self.keys),
                    callsets=[cs for _, cs in order])


class RoutedModelShard(ModelShard):
    """Peer routing (nb_route_*): a pair whose key another rank owns never enters this rank's table — k_pair appends its
    record to the owner's inbox.  The NVLink peer stores are modelled by an object all_gather when the inbox is merged."""

    def __init__(self, names_sorted, rank, world):
        super().__init__(names_sorted)
        self.rank, self.world = rank, world
        self.outbox = [[] for _ in range(world)]

    def add_pair(self, key, order, insertable, callset):
        if not insertable:
            return
        lo = key & (2 ** 64 - 1)
        owner = ((lo >> 40) & 0xFFFF) % self.world
        if owner == self.rank:
            return super().add_pair(key, order, insertable, callset)
        tag = 0
        if callset is not None:
            tag = _tag(callset)
            self.cs[tag] = callset      # interned locally, travels with the dictionary exchange
        self.outbox[owner].append((key, order, tag))

    def keys_export_partitioned(self, world):
        raise AssertionError("a routed merge never exports its key table")

    def route_sent(self):
        return [len(b) for b in self.outbox]

    def route_import(self, counts):
        everyone = [None] * self.world
        dist.all_gather_object(everyone, self.outbox)
        n = 0
        for r in range(self.world):
            if r == self.rank:
                continue
            assert int(counts[r]) == len(everyone[r][self.rank])      # the counts the merge delivered are the records that arrived
            for key, order, tag in everyone[r][self.rank]:
                n += 1
                if key not in self.keys or self.keys[key][0] < order:
                    self.keys[key] = (order, tag)
        self.outbox = [[] for _ in range(self.world)]
        return n


def _inputs():
    L = synth.SynthLibrary(seed=77, n_fam=40, n_all=5, group_on="")
    obj = L.to_json_obj()
    r1, o1, r2, o2 = synth.pairs(L, 0, 6000, seed=77, dup_rate=0.3)
    return L, obj, (r1, o1, r2, o2)


def _per_pair(o, ocfg, r1, o1, r2, o2, lo, hi):
    """Oracle over pairs [lo, hi): per pair (read_key, insertable, callset names or None)."""
    a1, b1 = r1[int(o1[lo]):int(o1[hi])], (o1[lo:hi + 1] - o1[lo]).astype(np.uint64)
    a2, b2 = r2[int(o2[lo]):int(o2[hi])], (o2[lo:hi + 1] - o2[lo]).astype(np.uint64)
    ref = o.run(a1, b1, a2, b2, threads=2)
    out = []
    for p in range(hi - lo):
        s1 = bytes(a1[int(b1[p]):int(b1[p + 1])]); s2 = bytes(a2[int(b2[p]):int(b2[p + 1])])
        key = int.from_bytes(hashlib.blake2b(s1 + s2, digest_size=16).digest(), "little")       # read_key = R1 string + R2 string (src/align.rs:576-579)
        ins = bool(ref["read_pass"][2 * p] or ref["read_pass"][2 * p + 1])                       # reached score_map (src/align.rs:685)
        cs = tuple(ref["scopes"][0][ref["pair_callset"][p]][0]) if ref["pair_counted"][p] else None
        out.append((key, ins, cs))
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, obj, (r1, o1, r2, o2) = _inputs()
        ocfg, oref = orc.parse_reference_library(obj, "unstranded")
        o = orc.Oracle(ocfg, oref)
        n = len(o1) - 1; per = n // world; lo, hi = rank * per, (rank + 1) * per if rank + 1 < world else n
        pairs = _per_pair(o, ocfg, r1, o1, r2, o2, lo, hi)
        import nimble_aligner_b200.multigpu as mg
        mg._ROUTED_CAP[0] = 16          # far below the dictionary sizes here: the block capacity must ratchet up, alike on both ranks
        for routed in (False, True):
            shard = RoutedModelShard(sorted(L.names), rank, world) if routed else ModelShard(sorted(L.names))
            for i, (key, ins, cs) in enumerate(pairs):
                shard.add_pair(key, lo + i + 1, ins, cs)
            raw, uniq = merge_across_ranks(shard, torch, dist, rank, world, "cpu", routed=routed)
            if routed:
                assert mg._ROUTED_CAP[0] >= len(shard.cs) > 16
            if rank == 0:
                q.put(({cs: int(c) for cs, c in zip(raw["callsets"], raw["dense_counts"].tolist()) if c}, uniq))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_merge_equals_single_process_counts():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, uniq = q.get(timeout=240)
    merged_routed, uniq_routed = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert merged_routed == merged and uniq_routed == uniq   # peer-routed records give the same tables as the all-to-all
    # single process over the union
    L, obj, (r1, o1, r2, o2) = _inputs()
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    ref = orc.Oracle(ocfg, oref).run(r1, o1, r2, o2, threads=2)
    want = {tuple(cs): int(c) for cs, c in ref["scopes"][0]}
    assert merged == want
    keys = {k for k, ins, _ in _per_pair(orc.Oracle(ocfg, oref), ocfg, r1, o1, r2, o2, 0, len(o1) - 1) if ins}
    assert uniq == len(keys) and len(want) > 50 and sum(want.values()) > 1000


# ------------------------------------------------------------------ BAM mode: scopes shard, per-cell tables add
class ScopedModelShard(ModelShard):
    def __init__(self, names_sorted):
        super().__init__(names_sorted)
        self.table = {}

    def add_scope(self, cell, rows):
        for cs, c in rows:
            cs = tuple(cs); self.cs[_tag(cs)] = cs
            self.table[(cell, cs)] = self.table.get((cell, cs), 0) + c

    def finalize(self):
        order = sorted(self.cs.items(), key=lambda kv: kv[1])
        dense = {cs: i for i, (_, cs) in enumerate(order)}
        items = sorted((cell, dense[cs], c) for (cell, cs), c in self.table.items())
        a = np.array(items, dtype=np.int64).reshape(-1, 3)
        return dict(callset_off=np.arange(len(order) + 1), row_scope=a[:, 0], row_callset=a[:, 1], row_count=a[:, 2], n_unique_keys=0, callsets=[cs for _, cs in order])


def _scoped_inputs():
    L = synth.SynthLibrary(seed=78, n_fam=40, n_all=5, group_on="")
    u = synth.umi_reads(L, 0, 1200, seed=9, n_cells=50)
    return L, L.to_json_obj(), u


def _scoped_rows(o, u, s0, s1):
    """Oracle over scopes [s0, s1): list of (cell, rows)."""
    start = np.concatenate([[0], np.cumsum(u["sizes"])]).astype(np.int64)
    a, b = int(start[s0]), int(start[s1]); L = 91
    bases, qual = u["bases"][a * L:b * L], u["qual"][a * L:b * L]
    off = (np.arange(b - a + 1, dtype=np.uint64) * L)
    scope_off = (start[s0:s1 + 1] - start[s0]).astype(np.uint64)
    ref = o.run(bases, off, bases, off, q1=qual, q2=qual, skip1=np.ones(b - a, dtype=np.uint8), skip2=None, scope_off=scope_off, threads=2, want_records=False)
    return [(int(u["cell"][start[s]]), ref["scopes"][s - s0]) for s in range(s0, s1)]


def _scoped_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, obj, u = _scoped_inputs()
        ocfg, oref = orc.parse_reference_library(obj, "unstranded")
        o = orc.Oracle(ocfg, oref)
        ns = len(u["sizes"]); per = ns // world; s0, s1 = rank * per, (rank + 1) * per if rank + 1 < world else ns
        shard = ScopedModelShard(sorted(L.names))
        for cell, rows in _scoped_rows(o, u, s0, s1):
            shard.add_scope(cell, rows)
        for force_sparse in (False, True):
            raw, cells, css, vals = merge_scoped_across_ranks(shard, torch, dist, rank, world, "cpu", 50 if not force_sparse else (1 << 40))
            if rank == 0:
                q.put({(int(c), raw["callsets"][int(k)]): int(v) for c, k, v in zip(cells, css, vals)})
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_scoped_merge_equals_single_process_table():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_scoped_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    dense, sparse = q.get(timeout=240), q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    L, obj, u = _scoped_inputs()
    ocfg, oref = orc.parse_reference_library(obj, "unstranded")
    want = {}
    for cell, rows in _scoped_rows(orc.Oracle(ocfg, oref), u, 0, len(u["sizes"])):
        for cs, c in rows:
            want[(cell, tuple(cs))] = want.get((cell, tuple(cs)), 0) + c
    assert dense == want and sparse == want and len(want) > 100
