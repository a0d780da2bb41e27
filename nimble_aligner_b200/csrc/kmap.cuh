// kmap.cuh — K2 k_map: length / entropy gates + seed-and-walk pseudo-alignment + colour intersection + thresholds.
// Included by kernels.cu after ReadView / EcAcc / cmp_fwd / cmp_bwd.
//
// Replaces align::pseudoalign (/root/reference/src/align.rs:945-989), Pseudoaligner::map_read_with_mismatch
// (call site src/align.rs:965; semantics SURVEY.md App. B) and filter_alignment_by_metrics (src/filter/align.rs:4-45).
//
// Execution model: persistent warps, two software stages per warp.
//   (S) seed stage, 32 fresh reads at a time at full lane efficiency: gates, two per-lane probes (the common hit), then
//       the whole warp searches the remaining stride-3 seeds of each still-seeking lane 32 at a time — first hit in
//       seed order wins, exactly the sequential search of App. B (an off-target read would otherwise hold its warp for
//       ~41 serial probes).  Seeded reads go into a per-warp shared-memory ring; gated / seedless reads are stored now.
//   (W) walk stage, one read per lane, one unitig per iteration: colour AND, base compare with the ordered per-node
//       mismatch budget, edge follow or re-seed.  A lane that finishes stores its read and pops the next seeded read
//       from the ring, so a long walk no longer leaves the other lanes idle.
#pragma once

// k-mer table lookup: bucketed cuckoo, exactly two 16-byte loads and four compares (no probe loop, no divergence)
__device__ __forceinline__ bool probe_kmer(const DevIndex& ix, const ReadView& rd, u32 pos, u32& node, u32& off) {
  u64 km = rd.win(pos) & KMASK;
  u32 b1, b2; nb_cuckoo_buckets(km, ix.tmask, b1, b2);
  const ulonglong2* T = (const ulonglong2*)ix.tkey;
  ulonglong2 k1 = __ldg(T + b1), k2 = __ldg(T + b2);
  u64 want = km | (1ULL << 63), slot;
  if (k1.x == want) slot = 2 * (u64)b1;
  else if (k1.y == want) slot = 2 * (u64)b1 + 1;
  else if (k2.x == want) slot = 2 * (u64)b2;
  else if (k2.y == want) slot = 2 * (u64)b2 + 1;
  else return false;
  u64 v = __ldg(ix.tval + slot); node = (u32)v; off = (u32)(v >> 32);
  return true;
}

// The warp searches the stride-3 seeds of lane l's read from position s_kp on; returns (to every lane) the first hit.
__device__ __forceinline__ bool coop_find(const DevIndex& ix, const BatchDev& b, u32 lane, u32 s_ri, u32 s_kp, u32 s_last, u32& f_kp, u32& f_node, u32& f_off, u32& tried) {
  const unsigned FULL = 0xFFFFFFFFu;
  ReadView srd{b.pk + (u64)s_ri * b.W, 1};
  tried = 0;
  for (u32 base = s_kp; base <= s_last; base += 96) {
    u32 my = base + 3 * lane, nd2 = 0, of2 = 0;
    bool hit = my <= s_last && probe_kmer(ix, srd, my, nd2, of2);
    unsigned hb = __ballot_sync(FULL, hit);
    if (hb) { int f = __ffs(hb) - 1; f_node = __shfl_sync(FULL, nd2, f); f_off = __shfl_sync(FULL, of2, f); f_kp = base + 3 * f; tried += f + 1; return true; }
    tried += min(32u, (s_last - base) / 3 + 1);   // same count as the sequential search: every seed of the round missed
  }
  return false;
}

enum { ST_DONE = 0, ST_SEED = 1, ST_WALK = 2 };
constexpr int RING = 64;   // seeded reads buffered per warp
constexpr int P_MIN = 6;   // idle lanes needed before the store/pop path runs
constexpr int S_MIN = 4;   // re-seeding lanes needed before the re-seed path runs

template <int COUNT_WORK>
__global__ void __launch_bounds__(128) k_map(BatchDev b, DevIndex ix, DevCfg cfg, Tables t) {
  __shared__ uint4 s_ring[4][RING];
  const unsigned FULL = 0xFFFFFFFFu;
  const u32 lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1;
  uint4* ring = s_ring[threadIdx.x >> 5];
  const u32* redge = (const u32*)ix.redge; const u32* ledge = (const u32*)ix.ledge;
  const u32 allowed = cfg.num_mismatches;
  WorkCnt wc = {0, 0, 0, 0};
  u32 r_head = 0, r_count = 0;                   // warp-uniform ring state
  bool drained = false;                          // warp-uniform: the global queue has no more reads for this warp
  // per-lane walk state
  int st = ST_DONE; bool has = false, first = true;
  u32 ri = 0, n = 0, cov = 0, mm = 0, kp = 0, node = 0, off = 0, last_kpos = 0;
  ReadView rd{b.pk, 1};
  EcAcc acc; acc.init(ix, t);
  for (;;) {
    // ---------------------------------------------------------------- (S) seed 32 fresh reads while the ring is low
    while (!drained && r_count < 32) {
      u32 base = 0;
      if (lane == 0) base = (u32)atomicAdd(&t.ctr->queue, 32ULL);
      base = __shfl_sync(FULL, base, 0);
      if (base + 32 >= b.n_reads) drained = true;
      if (COUNT_WORK && lane == 0) atomicAdd(&t.ctr->dbg[2], 1ULL);
      u32 q = base + lane; bool live = q < b.n_reads, seek = false, found = false;
      u32 qn = 0, qhdr = R_NO_MATCH, qkp = 0, qnode = 0, qoff = 0, qlast = 0;
      ReadView qrd{b.pk + (u64)(live ? q : 0) * b.W, 1};
      if (live) {
        u32 side = b.sides == 2 ? (q & 1) : 0; u64 p = b.sides == 2 ? (q >> 1) : q;
        qn = b.len_trim[q];
        bool skip = b.flags[side] != nullptr && (b.flags[side][p] & 1);
        if (skip) qhdr = R_SKIPPED | (1u << 10);                                        // src/align.rs:527-528
        else if (qn < cfg.min_read_len) qhdr = R_SHORT;                                 // src/align.rs:955-957
        else {
          // shannon_entropy on the (trimmed) read, src/utils.rs:96-119; terms come from a host-built table of
          // f*log2(f) (same libm as the CPU reference), summed in the reference's A,T,C,G order.
          u32 cC = 0, cG = 0, cT = 0;
          for (u32 w = 0; w * 32 < qn; w++) {
            u64 x = qrd.word(w); u32 c = min(32u, qn - w * 32);
            u64 lo = x & 0x5555555555555555ULL, hi = (x >> 1) & 0x5555555555555555ULL;
            u64 vm = c < 32 ? ((1ULL << (2 * c)) - 1) & 0x5555555555555555ULL : 0x5555555555555555ULL;
            cC += __popcll(lo & ~hi & vm); cG += __popcll(hi & ~lo & vm); cT += __popcll(hi & lo & vm);
          }
          u32 cA = qn - cC - cG - cT;
          const double* et = t.ent + (size_t)qn * (qn + 1) / 2;
          double e = 0.0;
          if (cA) e += et[cA];
          if (cT) e += et[cT];
          if (cC) e += et[cC];
          if (cG) e += et[cG];
          if (-e < 1.75) qhdr = R_ENTROPY;                                              // src/align.rs:960-962
          else if (qn >= (u32)K) { seek = true; qlast = qn - K; }                       // n < k: map_read returns None
        }
      }
      if (seek) {   // the seed at 0 and the next one, per lane
#pragma unroll 1
        for (int tries = 0; tries < 2 && !found && qkp <= qlast; tries++) {
          wc.probes++;
          if (probe_kmer(ix, qrd, qkp, qnode, qoff)) found = true; else qkp += 3;
        }
      }
      unsigned need = __ballot_sync(FULL, seek && !found && qkp <= qlast);
      while (need) {
        int l = __ffs(need) - 1; need &= need - 1;
        u32 s_kp = __shfl_sync(FULL, qkp, l), s_last = __shfl_sync(FULL, qlast, l), s_ri = __shfl_sync(FULL, q, l);
        u32 f_kp = 0, f_node = 0, f_off = 0, tried = 0;
        bool ok = coop_find(ix, b, lane, s_ri, s_kp, s_last, f_kp, f_node, f_off, tried);
        if ((int)lane == l) { wc.probes += tried; if (ok) { found = true; qkp = f_kp; qnode = f_node; qoff = f_off; } }
      }
      if (live && !found) {   // gated, or map_read_with_mismatch found no seed -> None -> NoMatch (src/align.rs:987)
        ReadRes rr; rr.hdr = qhdr; rr.score = 0; rr.mm = 0; rr.ec_len = 0; rr.bsize = 0; rr.ref = 0; rr.mask = 0;
        b.rres[q] = rr;
      }
      unsigned pm = __ballot_sync(FULL, found);
      if (found) ring[(r_head + r_count + __popc(pm & lt_mask)) % RING] = make_uint4(q, qkp, qnode, qoff);
      r_count += __popc(pm);
      __syncwarp();
    }
    // ---------------------------------------------------------------- (P) store finished reads, pop seeded ones.
    // Both side paths below run for whichever lanes need them; to keep them from executing at 1-4 active lanes on
    // every iteration they are batched: (P) waits for >= P_MIN idle lanes, (A)+(B) for >= S_MIN re-seeding lanes,
    // unless no lane could walk otherwise.
    unsigned walkers = __ballot_sync(FULL, st == ST_WALK);
    unsigned idle0 = __ballot_sync(FULL, st == ST_DONE);
    const bool do_pop = __popc(idle0) >= P_MIN || walkers == 0;
    if (do_pop && st == ST_DONE && has) {
      ReadRes rr; rr.hdr = R_NO_MATCH; rr.score = 0; rr.mm = 0; rr.ec_len = 0; rr.bsize = 0; rr.ref = 0; rr.mask = 0;
      if (acc.any) {
        u32 ecl = acc.ec_len();
        rr.score = (u16)cov; rr.mm = (u16)mm; rr.ec_len = ecl; rr.bsize = acc.big ? acc.alen : acc.bsize;
        rr.ref = acc.big ? acc.aoff : acc.boff; rr.mask = acc.mask;
        u32 reason;   // score as f64 / len as f64 >= score_percent  <=>  cov >= mincov[n] (host-built with the same division)
        if (cfg.discard_nonzero_mismatch && mm != 0) reason = R_NONZERO_MM;             // src/align.rs:971-973
        else if (cov >= cfg.score_threshold && cov >= (u32)__ldg(t.mincov + n) && ecl != 0) {   // src/filter/align.rs:17-45
          if (cfg.discard_multiple_matches && ecl > 1) reason = R_MULTI;
          else if (mm > cfg.num_mismatches) reason = R_ABOVE_MM;
          else reason = R_SUCCESS | (1u << 8);
        } else reason = R_SCORE_BELOW;
        rr.hdr = reason | (acc.big ? (1u << 9) : 0u);
      }
      b.rres[ri] = rr;
      has = false;
    }
    unsigned idle = idle0;
    if (do_pop && idle && r_count) {
      u32 npop = min((u32)__popc(idle), r_count), rank = __popc(idle & lt_mask);
      if (st == ST_DONE && rank < npop) {
        uint4 e = ring[(r_head + rank) % RING];
        ri = e.x; kp = e.y; node = e.z; off = e.w;
        n = b.len_trim[ri]; last_kpos = n - K;
        rd.p = b.pk + (u64)ri * b.W; rd.stride = 1;   // (staging the read in shared memory was measured: no gain, +10 % time)
        cov = 0; mm = 0; acc.reset(); first = true; has = true; st = ST_WALK;
      }
      r_head = (r_head + npop) % RING; r_count -= npop;
      __syncwarp();
    }
    if (__all_sync(FULL, st == ST_DONE)) { if (drained && r_count == 0) break; continue; }   // (do_pop was true: everything is stored)
    if (COUNT_WORK) {
      unsigned wl = __ballot_sync(FULL, st == ST_WALK), sl = __ballot_sync(FULL, st == ST_SEED);
      if (lane == 0) { atomicAdd(&t.ctr->dbg[0], 1ULL); atomicAdd(&t.ctr->dbg[1], (unsigned long long)__popc(wl)); atomicAdd(&t.ctr->dbg[3], (unsigned long long)__popc(sl)); atomicAdd(&t.ctr->dbg[4], (unsigned long long)r_count); if (drained) atomicAdd(&t.ctr->dbg[5], 1ULL); }
    }
    // ---------------------------------------------------------------- (A)+(B) re-seeding lanes (after a budget trip / dead end)
    unsigned seekers = __ballot_sync(FULL, st == ST_SEED);
    if (seekers && (__popc(seekers) >= S_MIN || __ballot_sync(FULL, st == ST_WALK) == 0)) {
      if (st == ST_SEED) {
#pragma unroll 1
        for (int tries = 0; tries < 2 && st == ST_SEED; tries++) {
          if (kp > last_kpos) { st = ST_DONE; break; }
          wc.probes++;
          if (probe_kmer(ix, rd, kp, node, off)) st = ST_WALK; else kp += 3;
        }
        if (st == ST_SEED && kp > last_kpos) st = ST_DONE;
      }
      unsigned need = __ballot_sync(FULL, st == ST_SEED);
      while (need) {
        int l = __ffs(need) - 1; need &= need - 1;
        u32 s_kp = __shfl_sync(FULL, kp, l), s_last = __shfl_sync(FULL, last_kpos, l), s_ri = __shfl_sync(FULL, ri, l);
        u32 f_kp = 0, f_node = 0, f_off = 0, tried = 0;
        bool ok = coop_find(ix, b, lane, s_ri, s_kp, s_last, f_kp, f_node, f_off, tried);
        if ((int)lane == l) { wc.probes += tried; if (ok) { kp = f_kp; node = f_node; off = f_off; st = ST_WALK; } else st = ST_DONE; }
      }
    }
    // ---------------------------------------------------------------- (C) left extension, only after the first seed and only
    //                                                                      if it sits at >= 20 % of the read [App. B]
    if (st == ST_WALK && first) {
      first = false;
      u32 lthr = (u32)(0.2 * (double)n);
      if (kp >= lthr) {
        u32 lp = kp - 1, pn = node, po = off > 0 ? off - 1 : 0;
        for (;;) {
          uint4 nd = __ldg(ix.node + pn);
          u64 start = (u64)nd.x | ((u64)(nd.w >> 8) << 32);
          u32 m = min(lp + 1, po + 1), mb, snp; bool brk;
          cmp_bwd(ix.unitig, start + po, rd, lp, m, allowed, mb, snp, brk);
          mm += snp; cov += mb; wc.bases += mb + (brk ? 1 : 0);
          if (lp + 1 - mb == 0 || brk) break;
          lp -= mb;
          u32 bs = rd.base(lp);
          if ((nd.w >> bs) & 1) {
            pn = __ldg(ledge + 4 * (u64)pn + bs);
            uint4 n2 = __ldg(ix.node + pn);
            po = n2.y - K; acc.add(n2.z, wc); wc.nodes++;
          } else break;
        }
      }
    }
    // ---------------------------------------------------------------- (D) one unitig of the forward walk
    if (st == ST_WALK) {
      uint4 nd = __ldg(ix.node + node);
      u64 start = (u64)nd.x | ((u64)(nd.w >> 8) << 32);
      kp += K; cov += K; acc.add(nd.z, wc); wc.nodes++;
      u32 ro = off + K, m = min(n - kp, nd.y - ro), mb, snp; bool brk;
      cmp_fwd(ix.unitig, start + ro, rd, kp, m, allowed, mb, snp, brk);
      mm += snp; cov += mb; kp += mb; wc.bases += mb + (brk ? 1 : 0);
      if (kp >= n) st = ST_DONE;
      else {
        u32 bs = rd.base(kp);
        if (!brk && ((nd.w >> (4 + bs)) & 1)) { node = __ldg(redge + 4 * (u64)node + bs); off = 0; kp -= K - 1; cov -= K - 1; }
        else st = kp > last_kpos ? ST_DONE : ST_SEED;
      }
    }
  }
  if (COUNT_WORK) {
    atomicAdd(&t.ctr->probes, (unsigned long long)wc.probes); atomicAdd(&t.ctr->nodes, (unsigned long long)wc.nodes);
    atomicAdd(&t.ctr->bases, (unsigned long long)wc.bases); atomicAdd(&t.ctr->colour_elems, (unsigned long long)wc.colour_elems);
  }
}
