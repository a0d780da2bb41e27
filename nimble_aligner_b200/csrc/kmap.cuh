// kmap.cuh — K2: length / entropy gates + seed-and-walk pseudo-alignment + colour intersection + thresholds, as two
// kernels: k_seed (gates + first seed of every read) and k_walk (the unitig walk).
// Included by kernels.cu after ReadView / EcAcc / cmp_bwd.
//
// Replaces align::pseudoalign (/root/reference/src/align.rs:945-989), Pseudoaligner::map_read_with_mismatch
// (call site src/align.rs:965; semantics SURVEY.md App. B) and filter_alignment_by_metrics (src/filter/align.rs:4-45).
//
// Execution model (round 2; what round 1's ncu captures showed is in DESIGN.md §4):
//   k_seed  one read per lane for the gates and the probe at position 0 (the common hit).  Reads still seeking — the
//           off-target reads, which must try all ~41 stride-3 seeds, and reads with an error in their first k-mer — go
//           to a shared-memory queue and the block searches them as a FLAT task list (entry, seed index): every lane
//           always has a probe to do, whatever mix of reads the block got (round 1 searched one read per warp round,
//           64 slots for 41 seeds, and spent 80 % of k_seed's instructions there).  First hit in seed order wins
//           (atomicMin on the seed index), exactly the sequential search of App. B.
//   k_walk  persistent warps, one seeded read per lane, ONE 32-BASE WINDOW per iteration: a lane enters a unitig
//           (64-byte walk record, colour AND) and compares its first window in the same iteration; longer stretches
//           continue window by window in the following iterations.  Every lane does the same bounded amount of work
//           per iteration, so the warp no longer waits in the longest lane's compare loop (round 1: 490 warp
//           instructions per iteration at 10 of 32 lanes inside that loop).
//   probes  the k-mer table is read through `probe_km`: blocked Bloom prefilter (one 64-bit word per query; L2-resident
//           with an evict_last hint when the table itself lives in HBM) in front of 32-byte buckets {key0, key1,
//           value0, value1}: a miss — 80 % of all probes — costs one word, a hit one bucket sector and needs no second
//           load for (unitig, offset).  The template flag BLOOM of the kernels means "HBM-resident index: use L2 hints".
#pragma once

#ifndef NB_P_MIN
#define NB_P_MIN 10
#endif
#ifndef NB_S_MIN
#define NB_S_MIN 4
#endif
constexpr int P_MIN = NB_P_MIN;   // idle lanes needed before the store/pop path runs
constexpr int S_MIN = NB_S_MIN;   // re-seeding lanes needed before the re-seed path runs

struct Bucket { u64 k0, k1, k2, k3; };
__device__ __forceinline__ Bucket ld_bucket(const u64* base, u64 b) {
  Bucket r;
  asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(base + 4 * b));
  return r;
}
// the same with an L2 eviction-priority hint: table buckets of an HBM-resident index are read once (evict_first) so that
// they do not push the prefilter (evict_last) out of L2
__device__ __forceinline__ Bucket ld_bucket_hint(const u64* base, u64 b, u64 policy) {
  Bucket r;
  asm("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(r.k0), "=l"(r.k1), "=l"(r.k2), "=l"(r.k3) : "l"(base + 4 * b), "l"(policy));
  return r;
}
__device__ __forceinline__ u64 ld_u64_hint(const u64* p, u64 policy) {
  u64 v;
  asm("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
  return v;
}
struct ProbePolicy { u64 first, last; };
template <int BLOOM> __device__ __forceinline__ ProbePolicy make_policy() {
  ProbePolicy p; p.first = p.last = 0;
  if (BLOOM) {
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.first));
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.last));
  }
  return p;
}

// exact membership of a 30-mer (km: first base in the low bits) and its (unitig, offset).  khash.h: h picks the bucket
// (multiply-high), g the Bloom word; the Bloom bits come from h's low bits, which the bucket choice hardly uses.
// HINTS: the index is HBM-resident (evict_first on table buckets, evict_last on the filter); FILTER: ask the Bloom word first
// (probes that mostly miss: seed searches; the probe at position 0 of a read mostly hits and goes straight to its bucket).
template <int HINTS, int FILTER>
__device__ __forceinline__ bool probe_km(const DevIndex& ix, const ProbePolicy& pol, u64 km, u32& node, u32& off) {
  const u64 hh = nb_khash(km);
  const u32 h = (u32)hh;
  if (FILTER) {
    const u64* wp = ix.bloom + __umulhi((u32)(hh >> 32), ix.bloom_words);
    const u64 w = HINTS ? ld_u64_hint(wp, pol.last) : __ldg(wp);
    const u32 nlo = nb_bloom_lo(h, ix.bloom_k), nhi = nb_bloom_hi(h);
    if (((u32)w & nlo) != nlo || ((u32)(w >> 32) & nhi) != nhi) return false;
  }
  u32 b = __umulhi(h, ix.n_pbuckets);
  const u64 want = km | (1ULL << 63);
  for (;;) {
    const Bucket k = HINTS ? ld_bucket_hint(ix.ptab, b, pol.first) : ld_bucket(ix.ptab, b);
    if (k.k0 == want) { node = (u32)k.k2; off = (u32)(k.k2 >> 32); return true; }
    if (k.k1 == want) { node = (u32)k.k3; off = (u32)(k.k3 >> 32); return true; }
    if (k.k1 == 0) return false;           // occupied slots are a prefix: the bucket has room, so the key is nowhere
    if (++b == ix.n_pbuckets) b = 0;       // 19 % of home buckets are full at the build load: the filter keeps misses away from this loop
  }
}
template <int HINTS>
__device__ __forceinline__ bool probe_kmer(const DevIndex& ix, const ProbePolicy& pol, const ReadView& rd, u32 pos, u32& node, u32& off) {
  return probe_km<HINTS, 1>(ix, pol, rd.win(pos) & KMASK, node, off);
}

// The warp searches the stride-3 seeds of lane l's read from position s_kp on, 32 per round; returns (to every lane) the
// first hit in seed order and the number of seeds a sequential search would have tried.  Only used by k_walk's re-seed
// path after the two per-lane probes missed (rare: the k-mer after a sequencing error usually hits at the second probe).
template <int BLOOM>
__device__ __forceinline__ bool coop_find(const DevIndex& ix, const ProbePolicy& pol, const BatchDev& b, u32 lane, u32 s_ri, u32 s_kp, u32 s_last, u32& f_kp, u32& f_node, u32& f_off, u32& tried) {
  const unsigned FULL = 0xFFFFFFFFu;
  ReadView srd{b.pk + (u64)s_ri * b.W, 1};
  tried = 0;
  for (u32 base = s_kp; base <= s_last; base += 96) {
    u32 my = base + 3 * lane, nd2 = 0, of2 = 0;
    bool hit = my <= s_last && probe_kmer<BLOOM>(ix, pol, srd, my, nd2, of2);
    unsigned hb = __ballot_sync(FULL, hit);
    if (hb) { int f = __ffs(hb) - 1; f_node = __shfl_sync(FULL, nd2, f); f_off = __shfl_sync(FULL, of2, f); f_kp = base + 3 * f; tried += f + 1; return true; }
    tried += min(32u, (s_last - base) / 3 + 1);   // same count as the sequential search: every seed of the round missed
  }
  return false;
}

enum { ST_DONE = 0, ST_SEED = 1, ST_WALK = 2 };

// ---- k_seed: 128 reads per block.  Gated / seedless reads get their final record here, seeded reads go to the list.
#ifndef NB_SEED_BLOCK
#define NB_SEED_BLOCK 128
#endif
constexpr int SEED_BLOCK = NB_SEED_BLOCK;
constexpr int SEED_R0 = 16;    // seeds per queue entry in the first cooperative round (an error in the first k-mer is passed within 10)
constexpr int SEED_R1 = 32;    // ... in the following rounds (off-target reads: all remaining seeds)
template <int COUNT_WORK, int BLOOM>
__global__ void __launch_bounds__(SEED_BLOCK) k_seed(BatchDev b, DevIndex ix, DevCfg cfg, Tables t) {
  const unsigned FULL = 0xFFFFFFFFu;
  __shared__ u32 q_ri[2][SEED_BLOCK], q_kp[2][SEED_BLOCK], q_last[2][SEED_BLOCK], q_best[SEED_BLOCK];
  __shared__ u32 q_n[2];
  const u32 tid = threadIdx.x, lane = tid & 31, lt_mask = (1u << lane) - 1;
  const ProbePolicy pol = make_policy<BLOOM>();
  u32 probes = 0;
  if (tid == 0) { q_n[0] = 0; q_n[1] = 0; }
  __syncthreads();
  // ---------------------------------------------------------------- per lane: gates, probe at position 0
  const u32 q = blockIdx.x * SEED_BLOCK + tid;
  const bool live = q < b.n_reads;
  bool seek = false, found = false;
  u32 qn = 0, qhdr = R_NO_MATCH, qnode = 0, qoff = 0, qlast = 0;
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
      // C = 01, G = 10, T = 11: popc(low bits) = C + T, popc(high bits) = G + T, popc(both) = T; bases behind the (trimmed)
      // end are cleared first and so read as A, which is counted from the length
      u32 pl = 0, ph = 0, cT = 0;
      for (u32 w = 0; w * 32 < qn; w++) {
        u64 x = qrd.word(w); const u32 c = qn - w * 32;
        if (c < 32) x &= (1ULL << (2 * c)) - 1;
        const u64 lo = x & 0x5555555555555555ULL, hi = (x >> 1) & 0x5555555555555555ULL;
        pl += __popcll(lo); ph += __popcll(hi); cT += __popcll(lo & hi);
      }
      const u32 cC = pl - cT, cG = ph - cT, cA = qn - pl - ph + cT;
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
  if (seek) {
    probes++;
    found = probe_km<BLOOM, 0>(ix, pol, qrd.word(0) & KMASK, qnode, qoff);
    if (!found && qlast >= 3) {       // still seeking: queue (read, next seed position, last seed position)
      u32 e = atomicAdd(&q_n[0], 1u);
      q_ri[0][e] = q; q_kp[0][e] = 3; q_last[0][e] = qlast;
    }
  }
  if (live && (!seek || (!found && qlast < 3))) {   // gated, or map_read_with_mismatch has no seed to find -> None -> NoMatch (src/align.rs:987)
    ReadRes rr; rr.hdr = qhdr; rr.score = 0; rr.mm = 0; rr.ec_len = 0; rr.bsize = 0; rr.ref = 0; rr.mask = 0;
    b.rres[q] = rr;
  }
  {
    unsigned pm = __ballot_sync(FULL, found);
    if (pm) {
      u32 at = 0;
      if (lane == 0) at = (u32)atomicAdd(&t.ctr->seeded_n, (unsigned long long)__popc(pm));
      at = __shfl_sync(FULL, at, 0);
      if (found) b.seeded[at + __popc(pm & lt_mask)] = make_uint4(q, 0u, qnode, qoff);
    }
  }
  // ---------------------------------------------------------------- block: flat (entry, seed) task list, round by round
  int cur = 0;
  for (u32 S = SEED_R0, lg = 4;; S = SEED_R1, lg = 5) {
    __syncthreads();
    const u32 E = q_n[cur];
    if (E == 0) break;
    for (u32 e = tid; e < E; e += SEED_BLOCK) q_best[e] = NONE32;
    if (tid == 0) q_n[cur ^ 1] = 0;
    __syncthreads();
    // two tasks per thread and iteration, their prefilter words requested together: a probe is a chain read words -> hash ->
    // filter word -> (rarely) bucket, and the kernel waits on that chain (ncu: 12 cycles of long-scoreboard stall per issue)
    for (u32 task = tid; task < (E << lg); task += 2 * SEED_BLOCK) {
      u64 km[2], hh[2], fw[2]; bool live2[2]; u32 ee[2], jj[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const u32 tk = task + u * SEED_BLOCK;
        live2[u] = tk < (E << lg);
        const u32 e = live2[u] ? tk >> lg : 0u, j = tk & (S - 1);
        ee[u] = e; jj[u] = j;
        const u32 pos = q_kp[cur][e] + 3 * j;
        live2[u] = live2[u] && pos <= q_last[cur][e] && j < q_best[e];      // (a hit at a smaller seed index already settles this entry)
        km[u] = 0; hh[u] = 0; fw[u] = 0;
        if (live2[u]) {
          ReadView srd{b.pk + (u64)q_ri[cur][e] * b.W, 1};
          km[u] = srd.win(pos) & KMASK; hh[u] = nb_khash(km[u]);
          const u64* wp = ix.bloom + __umulhi((u32)(hh[u] >> 32), ix.bloom_words);
          fw[u] = BLOOM ? ld_u64_hint(wp, pol.last) : __ldg(wp);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        if (!live2[u]) continue;
        const u32 h = (u32)hh[u], nlo = nb_bloom_lo(h, ix.bloom_k), nhi = nb_bloom_hi(h);
        if (((u32)fw[u] & nlo) != nlo || ((u32)(fw[u] >> 32) & nhi) != nhi) continue;
        u32 nd, of;
        if (probe_km<BLOOM, 0>(ix, pol, km[u], nd, of)) atomicMin(&q_best[ee[u]], jj[u]);
      }
    }
    __syncthreads();
    // one thread per entry: hit -> seeded list (its probe is repeated to fetch (unitig, offset): one per resolved read);
    // all seeds tried -> NoMatch; else the entry moves to the next round's queue
    for (u32 e0 = 0; e0 < E; e0 += SEED_BLOCK) {
      const u32 e = e0 + tid;
      bool hit = false; uint4 rec = make_uint4(0, 0, 0, 0);
      if (e < E) {
        const u32 ri = q_ri[cur][e], kp0 = q_kp[cur][e], last = q_last[cur][e], best = q_best[e];
        if (best != NONE32) {
          ReadView srd{b.pk + (u64)ri * b.W, 1};
          u32 nd = 0, of = 0;
          probe_kmer<BLOOM>(ix, pol, srd, kp0 + 3 * best, nd, of);
          hit = true; rec = make_uint4(ri, kp0 + 3 * best, nd, of);
          probes += best + 1;                                    // what the sequential search would have tried
        } else {
          probes += min(S, (last - kp0) / 3 + 1);                // the seeds of this round that exist, all missed
          const u32 nkp = kp0 + 3 * S;
          if (nkp <= last) {
            u32 ne = atomicAdd(&q_n[cur ^ 1], 1u);
            q_ri[cur ^ 1][ne] = ri; q_kp[cur ^ 1][ne] = nkp; q_last[cur ^ 1][ne] = last;
          } else {
            ReadRes rr; rr.hdr = R_NO_MATCH; rr.score = 0; rr.mm = 0; rr.ec_len = 0; rr.bsize = 0; rr.ref = 0; rr.mask = 0;
            b.rres[ri] = rr;
          }
        }
      }
      unsigned pm = __ballot_sync(FULL, hit);
      if (pm) {
        u32 at = 0;
        if (lane == 0) at = (u32)atomicAdd(&t.ctr->seeded_n, (unsigned long long)__popc(pm));
        at = __shfl_sync(FULL, at, 0);
        if (hit) b.seeded[at + __popc(pm & lt_mask)] = rec;
      }
    }
    cur ^= 1;
  }
  if (COUNT_WORK) {
    u32 p = probes;
    for (int o = 16; o; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
    if (lane == 0 && p) atomicAdd(&t.ctr->probes, (unsigned long long)p);
  }
}

// ---- k_walk: persistent warps, one seeded read per lane, one 32-base window per iteration
#ifndef NB_WALK_MINB
#define NB_WALK_MINB 8
#endif
template <int COUNT_WORK, int BLOOM>
__global__ void __launch_bounds__(128, NB_WALK_MINB) k_walk(BatchDev b, DevIndex ix, DevCfg cfg, Tables t) {
  const unsigned FULL = 0xFFFFFFFFu;
  const u32 lane = threadIdx.x & 31, lt_mask = (1u << lane) - 1;
  const u32* ledge = (const u32*)ix.ledge;
  const u32 allowed = cfg.num_mismatches;
  const u32 total = (u32)t.ctr->seeded_n;        // written by k_seed, complete at this kernel's start
  const ProbePolicy pol = make_policy<BLOOM>();
  WorkCnt wc = {0, 0, 0, 0};
  bool drained = false;                          // warp-uniform: the seeded list has no more reads for this warp
  int st = ST_DONE; bool has = false, first = true, enter = true;
  u32 ri = 0, n = 0, cov = 0, mm = 0, kp = 0, node = 0, off = 0, last_kpos = 0;
  u32 m = 0, snp = 0;                            // current unitig: bases left to compare, mismatches spent of its budget
  u32 nexts = 0, start_lo = 0; u64 e01 = 0, e23 = 0;   // current unitig: exts | start_hi<<8, start, right edges (A,C | G,T)
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
        rr.hdr = reason | (acc.big ? (1u << 9) : 0u) | (acc.uni ? (1u << 11) : 0u);
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
        rd.p = b.pk + (u64)ri * b.W; rd.stride = 1;
        cov = 0; mm = 0; acc.reset(); first = true; enter = true; has = true; st = ST_WALK;
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
          if (probe_kmer<BLOOM>(ix, pol, rd, kp, node, off)) { st = ST_WALK; enter = true; } else kp += 3;
        }
        if (st == ST_SEED && kp > last_kpos) st = ST_DONE;
      }
      unsigned need = __ballot_sync(FULL, st == ST_SEED);
      while (need) {
        int l = __ffs(need) - 1; need &= need - 1;
        u32 s_kp = __shfl_sync(FULL, kp, l), s_last = __shfl_sync(FULL, last_kpos, l), s_ri = __shfl_sync(FULL, ri, l);
        u32 f_kp = 0, f_node = 0, f_off = 0, tried = 0;
        bool ok = coop_find<BLOOM>(ix, pol, b, lane, s_ri, s_kp, s_last, f_kp, f_node, f_off, tried);
        if ((int)lane == l) { wc.probes += tried; if (ok) { kp = f_kp; node = f_node; off = f_off; st = ST_WALK; enter = true; } else st = ST_DONE; }
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
          u32 mq = min(lp + 1, po + 1), mb, sn; bool brk;
          cmp_bwd(ix.unitig, start + po, rd, lp, mq, allowed, mb, sn, brk);
          mm += sn; cov += mb; wc.bases += mb + (brk ? 1 : 0);
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
    // ---------------------------------------------------------------- (D) one window of the forward walk.  Entering a unitig
    // reads its 64-byte walk record (node fields, the four right edges, the colour's bitmap metadata and the first 64
    // bases: two 256-bit loads of one line), so a unitig whose compared stretch ends within its first 64 bases — most
    // of them — touches nothing else of the index.
    __syncwarp();   // lanes coming out of the re-seed / left-extension paths rejoin the walkers here: without it the compiler lets them run the window code below a second time on their own
    if (st == ST_WALK) {
      // Every step reads the record's second sector — colour bitmap metadata and the unitig's bases [K, K + 64): a forward
      // compare never starts below base K — so a stretch that continues into a second window finds its bases where the first
      // window found them (an L2 hit) instead of in the unitig array (three scattered loads on a divergent path in 86 % of
      // the iterations, and a DRAM burst each when the index lives in HBM).
      const u64* wp = (const u64*)(ix.walk + 4 * (u64)node);
      const Bucket fb = ld_bucket(wp, 1);
      const u64 s0 = fb.k2, s1 = fb.k3;
      if (enter) {
        const Bucket fa = ld_bucket(wp, 0);
        const u32 nlen = (u32)(fa.k0 >> 32);
        start_lo = (u32)fa.k0; nexts = (u32)(fa.k1 >> 32); e01 = fa.k2; e23 = fa.k3;
        kp += K; cov += K;
        acc.add((u32)fa.k1, make_uint4((u32)fb.k0, (u32)(fb.k0 >> 32), (u32)fb.k1, (u32)(fb.k1 >> 32)), wc); wc.nodes++;
        off += K; m = min(n - kp, nlen - off); snp = 0;
        enter = false;
      }
      const u32 o2 = off - K;                    // record base j = unitig base K + j (off >= K from the first step on)
      // the read's bases [kp, kp + 64): the window to compare and, right behind it, the base that picks the next edge (the
      // row's zero pad word makes word w + 1 always readable), so a step never waits for a second, dependent read load
      const u32 rsh = (kp & 31) * 2;
      const u64 rlo = rd.word(kp >> 5), rhi = rd.word((kp >> 5) + 1);
      const u64 rw = rsh ? (rlo >> rsh) | (rhi << (64 - rsh)) : rlo;
      // compare read[kp, kp+c) with unitig[off, off+c): the (allowed+1)-th mismatch of this unitig trips its budget — it is
      // counted in the mismatches but not in the coverage [App. B]
      const u32 c = min(32u, m);
      bool brk = false;
      if (c) {
        u64 uw;
        if (o2 + c <= 64) { if (o2 < 32) { u32 sh = 2 * o2; uw = sh ? (s0 >> sh) | (s1 << (64 - sh)) : s0; } else uw = s1 >> (2 * (o2 - 32)); }
        else uw = uwin3(ix.unitig, ((u64)start_lo | ((u64)(nexts >> 8) << 32)) + off);
        const u64 x = uw ^ rw;
        u64 d = (x | (x >> 1)) & 0x5555555555555555ULL;
        if (c < 32) d &= (1ULL << (2 * c)) - 1;
        const u32 cnt = (u32)__popcll(d);
        u32 adv = c;
        if (snp + cnt <= allowed) { snp += cnt; mm += cnt; }
        else {
          for (u32 i = snp; i < allowed; i++) d &= d - 1;
          adv = (u32)(__ffsll((long long)d) - 1) >> 1;
          mm += allowed + 1 - snp; brk = true;
        }
        kp += adv; cov += adv; off += adv; m -= adv; wc.bases += adv + (brk ? 1 : 0);
      }
      if (brk || m == 0) {                       // this unitig is finished: follow the read's next base or re-seed
        enter = true;
        if (kp >= n) st = ST_DONE;
        else {
          const u32 q = (rsh >> 1) + c;          // where read base kp (after the advance) sits in (rlo, rhi): <= 63
          const u32 bs = (u32)((q < 32 ? rlo >> (2 * q) : rhi >> (2 * (q - 32))) & 3u);
          if (!brk && ((nexts >> (4 + bs)) & 1)) { const u64 e = bs & 2 ? e23 : e01; node = bs & 1 ? (u32)(e >> 32) : (u32)e; off = 0; kp -= K - 1; cov -= K - 1; }
          else st = kp > last_kpos ? ST_DONE : ST_SEED;
        }
      }
    }
  }
  if (COUNT_WORK) {
    atomicAdd(&t.ctr->probes, (unsigned long long)wc.probes); atomicAdd(&t.ctr->nodes, (unsigned long long)wc.nodes);
    atomicAdd(&t.ctr->bases, (unsigned long long)wc.bases); atomicAdd(&t.ctr->colour_elems, (unsigned long long)wc.colour_elems);
  }
}
