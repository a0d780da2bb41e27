// kmap.cuh — K2: length / entropy gates + seed-and-walk pseudo-alignment + colour intersection + thresholds, as two
// kernels: k_seed (gates + first seed of every read) and k_walk (the unitig walk).
// Included by kernels.cu after ReadView / EcAcc / cmp_fwd / cmp_bwd.
//
// Replaces align::pseudoalign (/root/reference/src/align.rs:945-989), Pseudoaligner::map_read_with_mismatch
// (call site src/align.rs:965; semantics SURVEY.md App. B) and filter_alignment_by_metrics (src/filter/align.rs:4-45).
//
// Execution model:
//   k_seed  one read per lane, 32 consecutive reads per warp, at full lane efficiency: gates, two per-lane probes (the
//           common hit), then the whole warp searches the remaining stride-3 seeds of each still-seeking lane 64 at a
//           time — first hit in seed order wins, exactly the sequential search of App. B (an off-target read would
//           otherwise hold its warp for ~41 serial probes).  Gated / seedless reads get their final record here; seeded
//           reads go to a global list {read, seed position, node, offset} (16 B per read).  Finding a seed needs almost
//           no per-lane state (40 registers), so this half runs at 48 warps per SM.
//   k_walk  persistent warps, one seeded read per lane, one unitig per iteration: colour AND, base compare with the
//           ordered per-node mismatch budget, edge follow or re-seed.  A lane that finishes stores its read and pops the
//           next one from the list (batched: the pop runs when >= P_MIN lanes are idle).
// A fused single-kernel form (seed stage feeding a shared-memory ring inside the walk loop) gave the same results
// 4-7 % slower: it carried the walk state through the seed search (64 registers, 32 warps per SM).
#pragma once

// k-mer table lookup (khash.h): one 256-bit load fetches the four keys of the home bucket = one 32-byte sector; a
// bucket that neither holds the key nor has an empty slot (occupied slots are a prefix, so that is keys[3] != 0)
// sends the search to the next bucket.  A miss costs 1.07 sectors on average, a hit one more for (node, offset).
struct Bucket { u64 k0, k1, k2, k3; };
__device__ __forceinline__ Bucket ld_bucket(const u64* tkey, u64 b) {
  Bucket r;
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(tkey + 4 * b));
  return r;
}
// slot of `want` given its (already loaded) home bucket, or ~0
__device__ __forceinline__ u64 probe_finish(const DevIndex& ix, u64 want, u64 b, Bucket k) {
  for (;;) {   // (a branch-free select form of these compares was measured: +6 % k_seed time)
#ifdef NB_PROBE_LO32   // unmeasured variant: 80 % of probes miss, so reject the bucket on the low 32 bits of its keys (four compares accumulated into one predicate) before the 64-bit compare-and-branch chain
    const u32 wl = (u32)want;
    if (!(((u32)k.k0 == wl) | ((u32)k.k1 == wl) | ((u32)k.k2 == wl) | ((u32)k.k3 == wl))) {
      if (k.k3 == 0) return ~0ULL;
      if (++b == ix.n_buckets) b = 0;
      k = ld_bucket(ix.tkey, b);
      continue;
    }
#endif
    if (k.k0 == want) return 4 * b;
    if (k.k1 == want) return 4 * b + 1;
    if (k.k2 == want) return 4 * b + 2;
    if (k.k3 == want) return 4 * b + 3;
    if (k.k3 == 0) return ~0ULL;
    if (++b == ix.n_buckets) b = 0;   // rare (7 % of buckets are full at load 0.4)
    k = ld_bucket(ix.tkey, b);
  }
}
__device__ __forceinline__ bool probe_kmer(const DevIndex& ix, const ReadView& rd, u32 pos, u32& node, u32& off) {
  u64 km = rd.win(pos) & KMASK;
  u64 b = nb_table_bucket(km, ix.n_buckets);
  u64 slot = probe_finish(ix, km | (1ULL << 63), b, ld_bucket(ix.tkey, b));
  if (slot == ~0ULL) return false;
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
#ifndef NB_P_MIN
#define NB_P_MIN 6
#endif
#ifndef NB_S_MIN
#define NB_S_MIN 4
#endif
constexpr int P_MIN = NB_P_MIN;   // idle lanes needed before the store/pop path runs
constexpr int S_MIN = NB_S_MIN;   // re-seeding lanes needed before the re-seed path runs

// The warp searches the stride-3 seeds of one read from position s_kp on, 64 seeds per round (two per lane, the four
// bucket loads of a round in flight together: an off-target 150 bp read is settled in one round trip instead of two);
// returns (to every lane) the first hit in seed order and the number of seeds a sequential search would have tried.
__device__ __forceinline__ bool coop_find2(const DevIndex& ix, const BatchDev& b, u32 lane, u32 s_ri, u32 s_kp, u32 s_last, u32& f_kp, u32& f_node, u32& f_off, u32& tried) {
  const unsigned FULL = 0xFFFFFFFFu;
  ReadView srd{b.pk + (u64)s_ri * b.W, 1};
  tried = 0;
  for (u32 base = s_kp; base <= s_last; base += 192) {
    u32 p0 = base + 3 * lane, p1 = p0 + 96;
    bool v0 = p0 <= s_last, v1 = p1 <= s_last;
    u64 w0 = (srd.win(v0 ? p0 : s_last) & KMASK), w1 = (srd.win(v1 ? p1 : s_last) & KMASK);
    u64 b0 = nb_table_bucket(w0, ix.n_buckets), b1 = nb_table_bucket(w1, ix.n_buckets);
    Bucket ka = ld_bucket(ix.tkey, b0), kb = ld_bucket(ix.tkey, b1);
    u64 s0 = probe_finish(ix, w0 | (1ULL << 63), b0, ka), s1 = probe_finish(ix, w1 | (1ULL << 63), b1, kb);
    unsigned hb0 = __ballot_sync(FULL, v0 && s0 != ~0ULL), hb1 = __ballot_sync(FULL, v1 && s1 != ~0ULL);
    if (hb0 | hb1) {
      int f = hb0 ? __ffs(hb0) - 1 : __ffs(hb1) - 1;
      u32 nd2 = 0, of2 = 0;
      if ((int)lane == f) { u64 v = __ldg(ix.tval + (hb0 ? s0 : s1)); nd2 = (u32)v; of2 = (u32)(v >> 32); }
      f_node = __shfl_sync(FULL, nd2, f); f_off = __shfl_sync(FULL, of2, f);
      u32 idx = (hb0 ? 0u : 32u) + (u32)f;
      f_kp = base + 3 * idx; tried += idx + 1; return true;
    }
    tried += min(64u, (s_last - base) / 3 + 1);   // same count as the sequential search: every seed of the round missed
  }
  return false;
}

// Variant (-DNB_SEED_PAIRED, not the default): two seeking reads per round, each half-warp searches the stride-3 seeds of
// one read, 48 per round (three per lane, the bucket loads of both halves in flight together), so the serial chain of
// round trips a warp sits through — one per seeking lane with coop_find2 — halves; a 150 bp read (41 seeds) is settled in
// one round.  Bit-exact (166 parity / fuzz tests) but 56 registers instead of 40 (36 instead of 48 warps per SM) and
// the map stage went 0.539 -> 0.566 ms per 2 M reads: the seed stage is not bound by that chain.  `act`: this half has a read
// to search.  Results are uniform within a half: first hit in seed order and the number of seeds a sequential search
// would have tried.
__device__ __forceinline__ void coop_find_pair(const DevIndex& ix, const BatchDev& b, u32 lane, bool act, u32 s_ri, u32 s_kp, u32 s_last, bool& ok, u32& f_kp, u32& f_node, u32& f_off, u32& tried) {
  const unsigned FULL = 0xFFFFFFFFu;
  const u32 half = lane >> 4, hl = lane & 15;
  ReadView srd{b.pk + (u64)s_ri * b.W, 1};
  ok = false; tried = 0; f_kp = 0; f_node = 0; f_off = 0;
  u32 base = s_kp;
  bool live = act && base <= s_last;
  while (__any_sync(FULL, live)) {
    u32 p0 = base + 3 * hl, p1 = p0 + 48, p2 = p0 + 96;
    bool v0 = live && p0 <= s_last, v1 = live && p1 <= s_last, v2 = live && p2 <= s_last;
    u32 q0 = v0 ? p0 : s_last, q1 = v1 ? p1 : s_last, q2 = v2 ? p2 : s_last;   // (a dead half re-reads its own last window: harmless)
    u64 w0 = srd.win(q0) & KMASK, w1 = srd.win(q1) & KMASK, w2 = srd.win(q2) & KMASK;
    u64 b0 = nb_table_bucket(w0, ix.n_buckets), b1 = nb_table_bucket(w1, ix.n_buckets), b2 = nb_table_bucket(w2, ix.n_buckets);
    Bucket ka = ld_bucket(ix.tkey, b0), kb = ld_bucket(ix.tkey, b1), kc = ld_bucket(ix.tkey, b2);
    u64 s0 = probe_finish(ix, w0 | (1ULL << 63), b0, ka), s1 = probe_finish(ix, w1 | (1ULL << 63), b1, kb), s2 = probe_finish(ix, w2 | (1ULL << 63), b2, kc);
    unsigned h0 = (__ballot_sync(FULL, v0 && s0 != ~0ULL) >> (16 * half)) & 0xFFFFu;
    unsigned h1 = (__ballot_sync(FULL, v1 && s1 != ~0ULL) >> (16 * half)) & 0xFFFFu;
    unsigned h2 = (__ballot_sync(FULL, v2 && s2 != ~0ULL) >> (16 * half)) & 0xFFFFu;
    bool hit = live && (h0 | h1 | h2);
    u32 j = h0 ? 0u : (h1 ? 1u : 2u);
    int f = hit ? __ffs(h0 ? h0 : (h1 ? h1 : h2)) - 1 : 0;                       // winning lane within the half
    u32 nd2 = 0, of2 = 0;
    if (hit && (int)hl == f) { u64 v = __ldg(ix.tval + (j == 0 ? s0 : (j == 1 ? s1 : s2))); nd2 = (u32)v; of2 = (u32)(v >> 32); }
    u32 src = 16 * half + (u32)f;
    u32 rn = __shfl_sync(FULL, nd2, src), ro = __shfl_sync(FULL, of2, src);
    if (hit) { u32 idx = 16 * j + (u32)f; f_node = rn; f_off = ro; f_kp = base + 3 * idx; tried += idx + 1; ok = true; live = false; }
    else if (live) { tried += min(48u, (s_last - base) / 3 + 1); base += 144; live = base <= s_last; }
  }
}

// ---- k_seed: one read per lane, 32 consecutive reads per warp.  Gated / seedless reads get their final record here.
template <int COUNT_WORK>
__global__ void __launch_bounds__(128) k_seed(BatchDev b, DevIndex ix, DevCfg cfg, Tables t) {
  const unsigned FULL = 0xFFFFFFFFu;
  const u32 lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1;
  WorkCnt wc = {0, 0, 0, 0};
  const u32 base = (blockIdx.x * 128u + threadIdx.x) & ~31u;
  {
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
#ifndef NB_SEED_PAIRED   // (default; -DNB_SEED_PAIRED: two seeking reads per round, measured slower — DESIGN.md §4)
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
        bool ok = coop_find2(ix, b, lane, s_ri, s_kp, s_last, f_kp, f_node, f_off, tried);
        if ((int)lane == l) { wc.probes += tried; if (ok) { found = true; qkp = f_kp; qnode = f_node; qoff = f_off; } }
      }
#else
      if (seek) {   // the seeds at 0 and 3, per lane, both bucket loads in flight together (one round trip, not two, when the first misses)
        bool two = qlast >= 3;
        u64 w0 = qrd.win(0) & KMASK, w1 = qrd.win(two ? 3 : 0) & KMASK;
        u64 b0 = nb_table_bucket(w0, ix.n_buckets), b1 = nb_table_bucket(w1, ix.n_buckets);
        Bucket ka = ld_bucket(ix.tkey, b0), kb = ld_bucket(ix.tkey, b1);
        u64 s0 = probe_finish(ix, w0 | (1ULL << 63), b0, ka);
        wc.probes++;
        u64 slot = s0;
        if (s0 == ~0ULL) {
          qkp = 3;
          if (two) { wc.probes++; slot = probe_finish(ix, w1 | (1ULL << 63), b1, kb); if (slot == ~0ULL) qkp = 6; }
        }
        if (slot != ~0ULL) { u64 v = __ldg(ix.tval + slot); qnode = (u32)v; qoff = (u32)(v >> 32); found = true; }
      }
      unsigned need = __ballot_sync(FULL, seek && !found && qkp <= qlast);
      while (need) {   // two seeking lanes per round, one per half-warp
        int l0 = __ffs(need) - 1; need &= need - 1;
        int l1 = need ? __ffs(need) - 1 : -1; if (l1 >= 0) need &= need - 1;
        int mine = lane < 16 ? l0 : l1, src = mine < 0 ? l0 : mine;
        u32 s_kp = __shfl_sync(FULL, qkp, src), s_last = __shfl_sync(FULL, qlast, src), s_ri = __shfl_sync(FULL, q, src);
        u32 f_kp, f_node, f_off, tried; bool ok;
        coop_find_pair(ix, b, lane, mine >= 0, s_ri, s_kp, s_last, ok, f_kp, f_node, f_off, tried);
        u32 okb = __ballot_sync(FULL, ok);
#pragma unroll
        for (int h = 0; h < 2; h++) {   // hand each half's answer (uniform within the half) to the lane that asked
          int l = h ? l1 : l0;
          u32 rk = __shfl_sync(FULL, f_kp, 16 * h), rn = __shfl_sync(FULL, f_node, 16 * h), ro = __shfl_sync(FULL, f_off, 16 * h), rt = __shfl_sync(FULL, tried, 16 * h);
          if (l >= 0 && (int)lane == l) { wc.probes += rt; if ((okb >> (16 * h)) & 1u) { found = true; qkp = rk; qnode = rn; qoff = ro; } }
        }
      }
#endif
      if (live && !found) {   // gated, or map_read_with_mismatch found no seed -> None -> NoMatch (src/align.rs:987)
        ReadRes rr; rr.hdr = qhdr; rr.score = 0; rr.mm = 0; rr.ec_len = 0; rr.bsize = 0; rr.ref = 0; rr.mask = 0;
        b.rres[q] = rr;
      }
      unsigned pm = __ballot_sync(FULL, found);
      if (pm) {
        u32 at = 0;
        if (lane == 0) at = (u32)atomicAdd(&t.ctr->seeded_n, (unsigned long long)__popc(pm));
        at = __shfl_sync(FULL, at, 0);
        if (found) b.seeded[at + __popc(pm & lt_mask)] = make_uint4(q, qkp, qnode, qoff);
      }
  }
  if (COUNT_WORK) {
    u32 p = wc.probes;
    for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
    if (lane == 0 && p) atomicAdd(&t.ctr->probes, (unsigned long long)p);
  }
}

// ---- k_walk: persistent warps, one seeded read per lane, one unitig per iteration (stages P, A+B, C, D of kmap.cuh)
#ifndef NB_WALK_MINB
#define NB_WALK_MINB 9    // 56 registers: 36 warps per SM (measured with the 64-byte walk record: 9 blocks beat 8 and 10, which spills)
#endif
template <int COUNT_WORK>
__global__ void __launch_bounds__(128, NB_WALK_MINB) k_walk(BatchDev b, DevIndex ix, DevCfg cfg, Tables t) {
  const unsigned FULL = 0xFFFFFFFFu;
  const u32 lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1;
  const u32* ledge = (const u32*)ix.ledge;
  const u32 allowed = cfg.num_mismatches;
  const u32 total = (u32)t.ctr->seeded_n;        // written by k_seed, complete at this kernel's start
  WorkCnt wc = {0, 0, 0, 0};
  bool drained = false;                          // warp-uniform: the seeded list has no more reads for this warp
  int st = ST_DONE; bool has = false, first = true;
  u32 ri = 0, n = 0, cov = 0, mm = 0, kp = 0, node = 0, off = 0, last_kpos = 0;
  ReadView rd{b.pk, 1};
  EcAcc acc; acc.init(ix, t);
  for (;;) {
    // ---------------------------------------------------------------- (P) store finished reads, pop seeded ones (batched: >= P_MIN idle lanes)
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
    if (do_pop && idle0 && !drained) {
      u32 want = __popc(idle0), at = 0, rank = __popc(idle0 & lt_mask);
      if (lane == 0) at = (u32)atomicAdd(&t.ctr->wqueue, (unsigned long long)want);
      at = __shfl_sync(FULL, at, 0);
      if (at + want >= total) drained = true;
      if (st == ST_DONE && at + rank < total) {
        uint4 e = b.seeded[at + rank];
        ri = e.x; kp = e.y; node = e.z; off = e.w;
        n = b.len_trim[ri]; last_kpos = n - K;
        rd.p = b.pk + (u64)ri * b.W; rd.stride = 1;   // (copying the read into a shared-memory column was measured twice: -15 % L2 sectors, +17 % instructions, no gain)
        cov = 0; mm = 0; acc.reset(); first = true; has = true; st = ST_WALK;
      }
    }
    if (__all_sync(FULL, st == ST_DONE)) { if (drained) break; continue; }   // (do_pop was true: everything is stored)
    if (COUNT_WORK) {
      unsigned wl = __ballot_sync(FULL, st == ST_WALK), sl = __ballot_sync(FULL, st == ST_SEED);
      if (lane == 0) { atomicAdd(&t.ctr->dbg[0], 1ULL); atomicAdd(&t.ctr->dbg[1], (unsigned long long)__popc(wl)); atomicAdd(&t.ctr->dbg[3], (unsigned long long)__popc(sl)); if (drained) atomicAdd(&t.ctr->dbg[5], 1ULL); }
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
    // ---------------------------------------------------------------- (D) one unitig of the forward walk, off its 64-byte walk record:
    // node fields, the four right edges, the colour's bitmap metadata and the first 64 bases arrive together (two
    // 256-bit loads of one line), so a step on a unitig of <= 64 bases touches nothing else of the index
    if (st == ST_WALK) {
      const u64* wp = (const u64*)(ix.walk + 4 * (u64)node);
      Bucket fa = ld_bucket(wp, 0), fb = ld_bucket(wp, 1);
      const u32 nlen = (u32)(fa.k0 >> 32), ncol = (u32)fa.k1, nexts = (u32)(fa.k1 >> 32);
      const u64 start = (fa.k0 & 0xFFFFFFFFULL) | ((u64)(nexts >> 8) << 32);
      kp += K; cov += K;
      acc.add(ncol, make_uint4((u32)fb.k0, (u32)(fb.k0 >> 32), (u32)fb.k1, (u32)(fb.k1 >> 32)), wc); wc.nodes++;
      u32 ro = off + K, m = min(n - kp, nlen - ro), mb, snp, nxb; bool brk;
      cmp_fwd(ix.unitig, start, ro, fb.k2, fb.k3, rd, kp, m, allowed, mb, snp, brk, nxb);
      mm += snp; cov += mb; kp += mb; wc.bases += mb + (brk ? 1 : 0);
      if (kp >= n) st = ST_DONE;
      else {
        u32 bs = (!brk && nxb < 4) ? nxb : rd.base(kp);   // usually already in the compare's last window: one load less per unitig
        if (!brk && ((nexts >> (4 + bs)) & 1)) { u64 e = bs & 2 ? fa.k3 : fa.k2; node = bs & 1 ? (u32)(e >> 32) : (u32)e; off = 0; kp -= K - 1; cov -= K - 1; }
        else st = kp > last_kpos ? ST_DONE : ST_SEED;
      }
    }
  }
  if (COUNT_WORK) {
    atomicAdd(&t.ctr->probes, (unsigned long long)wc.probes); atomicAdd(&t.ctr->nodes, (unsigned long long)wc.nodes);
    atomicAdd(&t.ctr->bases, (unsigned long long)wc.bases); atomicAdd(&t.ctr->colour_elems, (unsigned long long)wc.colour_elems);
  }
}
