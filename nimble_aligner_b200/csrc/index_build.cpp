// index_build.cpp — host build of the coloured, compacted, stranded de Bruijn graph (k = 30) into the flat GPU layout.
//
// Replaces debruijn_mapping::build_index::build_index::<Kmer30>(seqs, names, {}, cores)
// (call sites /root/reference/src/bin/main.rs:121-128, tests/utils.rs:48-51; the crate itself is an un-vendored git
// dependency, Cargo.toml:23).  Semantics: every 30-mer of every sequence, stranded; colour =
// sorted set of sequence ids containing the k-mer; exts = bases observed before / after any occurrence; unitigs join
// x -> y iff |R(x)| == 1, |L(y)| == 1 and colour(x) == colour(y); a pure cycle starts at its smallest k-mer.
// Method here (not the upstream's minimizer shards + MPHF): enumerate occurrences, bucketed parallel sort, group,
// open-addressed bucketed table (khash.h) over the distinct k-mers, chain walk from the chain starts on threads.
#include <algorithm>
#include <atomic>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <unordered_map>

#include <unistd.h>
#include "host.hpp"
#include "khash.h"

using namespace nb;

namespace {

struct Occ { u64 kmer; u32 id; u32 lr; };  // kmer big-endian (first base most significant => lexicographic order); lr = l | r<<4, 4 = none

inline u64 mix64(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// big-endian 60-bit k-mer -> device form (base i at bits 2i)
inline u64 to_device_form(u64 be) {   // reverse the 32 two-bit groups of the word, then drop the two groups that were the unused top bits
  u64 x = __builtin_bswap64(be);
  x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
  x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
  return x >> (2 * (32 - K));
}

void parallel_for(int n_threads, u64 n, const std::function<void(u64, u64, int)>& fn) {
  if (n_threads <= 1 || n < 4096) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) { u64 a = n * (u64)t / (u64)n_threads, b = n * (u64)(t + 1) / (u64)n_threads; th.emplace_back(fn, a, b, t); }
  for (auto& x : th) x.join();
}

// k-mer table (khash.h): buckets of four keys, linear probing by bucket.
struct Table {
  const std::vector<u64>& ckey; std::vector<u64>* mkey; std::vector<u64>* mval; u64 nbuckets;
  // returns slot of kmer (device form) or ~0
  u64 find(u64 dk) const {
    u64 b = nb_table_bucket(dk, nbuckets), want = dk | (1ULL << 63);
    for (u64 tries = 0; tries < nbuckets; tries++) {
      for (u64 s = 4 * b; s < 4 * b + 4; s++) { if (ckey[s] == want) return s; if (!ckey[s]) return ~0ULL; }
      if (++b == nbuckets) b = 0;
    }
    return ~0ULL;
  }
  // distinct keys, load < 1: always finds room.  Safe from several threads at once: a slot is claimed by compare-and-swap
  // and slots of a bucket are only tried in order, so the occupied slots of a bucket stay a prefix and a key moves on to
  // the next bucket only when its home bucket is full — what find() and the device probe rely on.  Which of its
  // candidate slots a key ends up in depends on the interleaving (nb_index_compare looks keys up instead of comparing slots).
  void insert(u64 dk, u64 v) {
    u64* key = mkey->data(); u64* val = mval->data();
    u64 b = nb_table_bucket(dk, nbuckets), want = dk | (1ULL << 63);
    for (;;) {
      for (u64 s = 4 * b; s < 4 * b + 4; s++) {
        u64 expect = 0;
        if (__atomic_load_n(&key[s], __ATOMIC_RELAXED) == 0 && __atomic_compare_exchange_n(&key[s], &expect, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { val[s] = v; return; }
      }
      if (++b == nbuckets) b = 0;
    }
  }
};

}  // namespace

// Universes: connected components of sequences under "share a colour"; colours of components of <= 64 sequences get a
// 64-bit mask over the component's sorted member list (appended to col_ids).  Shared by the host and the GPU builder.
int nb_build_universes(nb_index* ix, u32 n_seq_arg) {
    u32 n_col = (u32)ix->col_off.size() - 1, n_seq = n_seq_arg, n_ids = ix->col_off[n_col];
    std::vector<u32> parent(n_seq);
    for (u32 i = 0; i < n_seq; i++) parent[i] = i;
    auto find = [&](u32 x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    for (u32 c = 0; c < n_col; c++) { u32 r0 = find(ix->col_ids[ix->col_off[c]]); for (u32 k = ix->col_off[c] + 1; k < ix->col_off[c + 1]; k++) { u32 r = find(ix->col_ids[k]); if (r != r0) parent[r] = r0; } }
    std::vector<u32> comp_size(n_seq, 0), comp_off(n_seq, NONE32), fill(n_seq, 0);
    for (u32 i = 0; i < n_seq; i++) comp_size[find(i)]++;
    std::vector<char> used(n_seq, 0);
    for (u32 c = 0; c < n_col; c++) used[find(ix->col_ids[ix->col_off[c]])] = 1;   // only components that own a colour
    u64 at = n_ids;
    for (u32 i = 0; i < n_seq; i++) if (used[i] && comp_size[i] <= 64) { comp_off[i] = (u32)at; at += comp_size[i]; }
    if (at >= 0xFFFFFFFFull) return fail(NB_ERR_UNSUPPORTED, "colour table exceeds 2^32 entries");
    ix->col_ids.resize(at);
    for (u32 i = 0; i < n_seq; i++) { u32 r = find(i); if (comp_off[r] != NONE32) ix->col_ids[comp_off[r] + fill[r]++] = i; }   // ascending ids
    ix->col_meta.assign(4 * (size_t)n_col, 0);
    for (u32 c = 0; c < n_col; c++) {
      u32 r = find(ix->col_ids[ix->col_off[c]]);
      if (comp_off[r] == NONE32) continue;
      const u32* u = &ix->col_ids[comp_off[r]]; u32 us = comp_size[r]; u64 mask = 0; u32 j = 0;
      for (u32 k = ix->col_off[c]; k < ix->col_off[c + 1]; k++) { while (u[j] != ix->col_ids[k]) j++; mask |= 1ULL << j; }
      ix->col_meta[4 * (size_t)c] = comp_off[r]; ix->col_meta[4 * (size_t)c + 1] = us; ix->col_meta[4 * (size_t)c + 2] = (u32)mask; ix->col_meta[4 * (size_t)c + 3] = (u32)(mask >> 32);
    }
    return NB_OK;
}

int nb_build_index_impl(const std::vector<std::vector<u8>>& seqs, int n_threads, nb_index** out) {
  if (n_threads < 1) n_threads = 1;
  if (seqs.size() >= 0xFFFFFFFFull) return fail(NB_ERR_UNSUPPORTED, "too many reference sequences");
  nb_index* ix = new nb_index();
  ix->n_sequences = seqs.size();
  static const bool stats = getenv("NB_INDEX_STATS") != nullptr;   // phase times on stderr
  auto now = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; };
  double t_prev = now();
  auto phase = [&](const char* name) { if (!stats) return; double t = now(); fprintf(stderr, "index build: %-28s %.2f s\n", name, t - t_prev); t_prev = t; };
  // ---- 1. enumerate occurrences
  std::vector<u64> occ_off(seqs.size() + 1, 0);
  for (size_t s = 0; s < seqs.size(); s++) occ_off[s + 1] = occ_off[s] + (seqs[s].size() >= (size_t)K ? seqs[s].size() - K + 1 : 0);
  u64 n_occ = occ_off.back();
  std::vector<Occ> occ(n_occ);
  parallel_for(n_threads, seqs.size(), [&](u64 a, u64 b, int) {
    for (u64 s = a; s < b; s++) {
      const std::vector<u8>& d = seqs[s];
      if (d.size() < (size_t)K) continue;
      u64 km = 0; Occ* o = &occ[occ_off[s]];
      for (size_t p = 0; p < d.size(); p++) {
        km = ((km << 2) | d[p]) & KMASK;
        if (p + 1 >= (size_t)K) {
          size_t st = p + 1 - K;
          o->kmer = km; o->id = (u32)s; o->lr = (st > 0 ? d[st - 1] : 4u) | ((p + 1 < d.size() ? d[p + 1] : 4u) << 4);
          o++;
        }
      }
    }
  });
  phase("enumerate");
  // ---- 2. bucketed parallel sort by (kmer, id): bucket = top 10 bits (first 5 bases)
  const int NB = 1024;
  std::vector<Occ> sorted(n_occ);
  std::vector<u64> bstart(NB + 1, 0);
  {
    std::vector<std::vector<u64>> cnt(n_threads, std::vector<u64>(NB, 0));
    parallel_for(n_threads, n_occ, [&](u64 a, u64 b, int t) { for (u64 i = a; i < b; i++) cnt[t][occ[i].kmer >> 50]++; });
    bool used_threads = !(n_threads <= 1 || n_occ < 4096);
    int T = used_threads ? n_threads : 1;
    std::vector<std::vector<u64>> pos(T, std::vector<u64>(NB, 0));
    u64 run = 0;
    for (int bkt = 0; bkt < NB; bkt++) { bstart[bkt] = run; for (int t = 0; t < T; t++) { pos[t][bkt] = run; run += cnt[t][bkt]; } }
    bstart[NB] = run;
    parallel_for(n_threads, n_occ, [&](u64 a, u64 b, int t) { for (u64 i = a; i < b; i++) sorted[pos[t][occ[i].kmer >> 50]++] = occ[i]; });
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int bkt = next.fetch_add(1); if (bkt >= NB) break; std::sort(sorted.begin() + bstart[bkt], sorted.begin() + bstart[bkt + 1], [](const Occ& x, const Occ& y) { return x.kmer != y.kmer ? x.kmer < y.kmer : x.id < y.id; }); } };
    std::vector<std::thread> th; for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work(); for (auto& x : th) x.join();
  }
  phase("sort");
  std::vector<Occ>().swap(occ);
  // ---- 3. group into distinct k-mers (sorted order): exts, colour signature; intern colours
  std::vector<u64> gstart;  // start offset of each distinct k-mer in `sorted`
  {
    std::vector<std::vector<u64>> parts(NB);
    std::atomic<int> next(0);
    auto work = [&]() { for (;;) { int bkt = next.fetch_add(1); if (bkt >= NB) break; auto& v = parts[bkt]; for (u64 i = bstart[bkt]; i < bstart[bkt + 1]; i++) if (i == bstart[bkt] || sorted[i].kmer != sorted[i - 1].kmer) v.push_back(i); } };
    std::vector<std::thread> th; for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work(); for (auto& x : th) x.join();
    u64 tot = 0; for (auto& v : parts) tot += v.size();
    gstart.reserve(tot + 1);
    for (auto& v : parts) gstart.insert(gstart.end(), v.begin(), v.end());
    gstart.push_back(n_occ);
  }
  u64 n = gstart.size() - 1;
  ix->n_kmers = n;
  std::vector<u64> kmers(n); std::vector<u8> L(n), R(n); std::vector<u32> col(n);
  std::vector<u64> sig_a(n), sig_b(n);
  parallel_for(n_threads, n, [&](u64 a, u64 b, int) {
    for (u64 g = a; g < b; g++) {
      u8 l = 0, r = 0; u64 ha = 0x9E3779B97F4A7C15ULL, hb = 0xC2B2AE3D27D4EB4FULL; u32 last = NONE32;
      for (u64 i = gstart[g]; i < gstart[g + 1]; i++) {
        u32 lr = sorted[i].lr;
        if ((lr & 15) < 4) l |= (u8)(1u << (lr & 15));
        if ((lr >> 4) < 4) r |= (u8)(1u << (lr >> 4));
        if (sorted[i].id != last) { last = sorted[i].id; ha = mix64(ha ^ last) + 0x632BE59BD9B4E019ULL; hb = mix64(hb + last * 0x9FB21C651E98DF25ULL); }
      }
      kmers[g] = sorted[gstart[g]].kmer; L[g] = l; R[g] = r; sig_a[g] = ha; sig_b[g] = hb;
    }
  });
  phase("group + signatures");
  {
    // colours are numbered in order of first appearance (ascending k-mer): every thread collects the first k-mer of each
    // signature in its range, the per-thread maps are merged (minimum first k-mer), signatures sorted by that k-mer get
    // their ids, and a second parallel pass labels the k-mers from the finished (read-only) map
    struct Sig { u64 a, b; bool operator==(const Sig& o) const { return a == o.a && b == o.b; } };
    struct SigHash { size_t operator()(const Sig& s) const { return (size_t)(s.a ^ (s.b * 0x9E3779B97F4A7C15ULL)); } };
    int T = (n_threads <= 1 || n < 4096) ? 1 : n_threads;
    std::vector<std::unordered_map<Sig, u64, SigHash>> local(T);
    for (auto& m : local) m.reserve(1 << 16);
    parallel_for(n_threads, n, [&](u64 a, u64 b, int t) {
      auto& m = local[t]; u64 pa = 0, pb = 0; bool have = false;
      for (u64 g = a; g < b; g++) {
        if (have && sig_a[g] == pa && sig_b[g] == pb) continue;
        pa = sig_a[g]; pb = sig_b[g]; have = true;
        Sig k{pa, pb}; if (m.find(k) == m.end()) m.emplace(k, g);   // keeps the first (smallest) g of the range; (emplace alone allocates a node before it looks)
      }
    });
    std::unordered_map<Sig, u64, SigHash> first;
    for (auto& m : local) { for (auto& kv : m) { auto it = first.find(kv.first); if (it == first.end()) first.emplace(kv.first, kv.second); else if (kv.second < it->second) it->second = kv.second; } m.clear(); }
    std::vector<std::pair<u64, Sig>> order; order.reserve(first.size());
    for (auto& kv : first) order.emplace_back(kv.second, kv.first);
    std::sort(order.begin(), order.end(), [](const std::pair<u64, Sig>& x, const std::pair<u64, Sig>& y) { return x.first < y.first; });
    std::unordered_map<Sig, u32, SigHash> intern; intern.reserve(order.size() * 2);
    ix->col_off.push_back(0);
    for (u32 c = 0; c < order.size(); c++) {
      intern.emplace(order[c].second, c);
      u64 g = order[c].first; u32 last = NONE32;
      for (u64 i = gstart[g]; i < gstart[g + 1]; i++) if (sorted[i].id != last) { last = sorted[i].id; ix->col_ids.push_back(last); }
      if (ix->col_ids.size() >= 0xFFFFFFFFull) { delete ix; return fail(NB_ERR_UNSUPPORTED, "colour table exceeds 2^32 entries"); }
      ix->col_off.push_back((u32)ix->col_ids.size());
    }
    parallel_for(n_threads, n, [&](u64 a, u64 b, int) {
      u64 pa = 0, pb = 0; u32 pc = NONE32;
      for (u64 g = a; g < b; g++) {
        if (pc != NONE32 && sig_a[g] == pa && sig_b[g] == pb) { col[g] = pc; continue; }
        pa = sig_a[g]; pb = sig_b[g]; pc = intern.find(Sig{pa, pb})->second; col[g] = pc;
      }
    });
  }
  phase("colour interning");
  { int urc = nb_build_universes(ix, (u32)seqs.size()); if (urc) { delete ix; return urc; } }
  std::vector<Occ>().swap(sorted); std::vector<u64>().swap(gstart); std::vector<u64>().swap(sig_a); std::vector<u64>().swap(sig_b);
  phase("universes");
  // ---- 4. bucketed cuckoo table over distinct k-mers (value = distinct index for now), load <= 0.5
  u64 slots = 0;
  {
    u64 nbk = nb_table_size(n);
    if (nbk > 0xFFFFFFFFull) { delete ix; return fail(NB_ERR_UNSUPPORTED, "k-mer table exceeds 2^34 slots"); }
    ix->table_buckets = nbk; slots = 4 * nbk;
    ix->table_key.assign(slots, 0); ix->table_val.assign(slots, 0);
    Table tb{ix->table_key, &ix->table_key, &ix->table_val, nbk};
    parallel_for(n_threads, n, [&](u64 a, u64 b, int) { for (u64 g = a; g < b; g++) tb.insert(to_device_form(kmers[g]), g); });
  }
  Table tab{ix->table_key, nullptr, nullptr, ix->table_buckets};
  auto index_of = [&](u64 be) -> u64 { u64 s = tab.find(to_device_form(be)); return s == ~0ULL ? ~0ULL : ix->table_val[s]; };
  phase("table insert");
  // ---- 5. join relation
  std::vector<u32> succ(n, NONE32), pred(n, NONE32);
  if (n >= 0xFFFFFFFFull) { delete ix; return fail(NB_ERR_UNSUPPORTED, "more than 2^32 distinct k-mers"); }
  parallel_for(n_threads, n, [&](u64 a, u64 b, int) {
    for (u64 g = a; g < b; g++) {
      if (__builtin_popcount(R[g]) != 1) continue;
      u64 y = ((kmers[g] << 2) | (u64)__builtin_ctz(R[g])) & KMASK;
      u64 j = index_of(y);
      if (j != ~0ULL && __builtin_popcount(L[j]) == 1 && col[g] == col[j]) { succ[g] = (u32)j; pred[j] = (u32)g; }  // pred[j] has a single writer: |L(j)| == 1
    }
  });
  phase("join");
  // ---- 6. unitigs: starts = k-mers without a joining predecessor; pure cycles start at their smallest k-mer
  std::vector<u64> starts;
  for (u64 g = 0; g < n; g++) if (pred[g] == NONE32) starts.push_back(g);
  std::vector<u32> node_of(n, NONE32), off_of(n, 0);
  std::vector<u32> node_len;  // in k-mers
  std::vector<u64> node_first, node_last;
  auto walk = [&](u64 s, u32 id) { u64 cur = s; u32 o = 0; for (;;) { node_of[cur] = id; off_of[cur] = o; u32 nx = succ[cur]; if (nx == NONE32 || node_of[nx] != NONE32) break; cur = nx; o++; } return std::make_pair(cur, o + 1); };
  node_len.resize(starts.size()); node_first.resize(starts.size()); node_last.resize(starts.size());
  parallel_for(n_threads, starts.size(), [&](u64 a, u64 b, int) { for (u64 i = a; i < b; i++) { auto r = walk(starts[i], (u32)i); node_first[i] = starts[i]; node_last[i] = r.first; node_len[i] = r.second; } });
  for (u64 g = 0; g < n; g++) if (node_of[g] == NONE32) {  // cycles (rare), ascending k-mer order
    u32 id = (u32)node_len.size(); auto r = walk(g, id); node_first.push_back(g); node_last.push_back(r.first); node_len.push_back(r.second);
  }
  phase("chain walks");
  u64 n_nodes = node_len.size();
  std::vector<u64> base_start(n_nodes + 1, 0);
  for (u64 i = 0; i < n_nodes; i++) base_start[i + 1] = base_start[i] + node_len[i] + K - 1;
  ix->unitig_bases = base_start[n_nodes];
  if (ix->unitig_bases >> 40) { delete ix; return fail(NB_ERR_UNSUPPORTED, "unitig store exceeds 2^40 bases"); }
  ix->unitig.assign((ix->unitig_bases + 31) / 32 + 2, 0);
  ix->node.resize(n_nodes); ix->redge.assign(4 * n_nodes, NONE32); ix->ledge.assign(4 * n_nodes, NONE32);
  parallel_for(n_threads, n_nodes, [&](u64 a, u64 b, int) {
    for (u64 i = a; i < b; i++) {
      u64 pos = base_start[i];
      auto put = [&](u64 base) { __atomic_fetch_or(&ix->unitig[pos >> 5], base << (2 * (pos & 31)), __ATOMIC_RELAXED); pos++; };
      u64 first = kmers[node_first[i]];
      for (int k = K - 1; k >= 0; k--) put((first >> (2 * k)) & 3);
      u64 cur = node_first[i];
      for (u32 o = 1; o < node_len[i]; o++) { cur = succ[cur]; put(kmers[cur] & 3); }
      NodeRec& nr = ix->node[i];
      nr.start_lo = (u32)base_start[i]; nr.len = node_len[i] + K - 1; nr.colour = col[node_first[i]];
      nr.exts_hi = (u32)L[node_first[i]] | ((u32)R[node_last[i]] << 4) | ((u32)(base_start[i] >> 32) << 8);
    }
  });
  phase("unitig bases + nodes");
  // table values -> (node, offset)
  parallel_for(n_threads, slots, [&](u64 a, u64 b, int) { for (u64 h = a; h < b; h++) if (ix->table_key[h] >> 63) { u64 g = ix->table_val[h]; ix->table_val[h] = (u64)node_of[g] | ((u64)off_of[g] << 32); } });
  // distinct index of a k-mer is gone from the table now; edges need start/end node of neighbouring k-mers, which the
  // table gives directly: right edge target must sit at offset 0, left edge target at its node's last k-mer.
  int bad = 0;
  parallel_for(n_threads, n_nodes, [&](u64 a, u64 b, int) {
    for (u64 i = a; i < b; i++) {
      u64 first = kmers[node_first[i]], last = kmers[node_last[i]];
      u8 l = L[node_first[i]], r = R[node_last[i]];
      for (int bb = 0; bb < 4; bb++) {
        if (r >> bb & 1) { u64 y = ((last << 2) | (u64)bb) & KMASK; u64 s = tab.find(to_device_form(y)); if (s == ~0ULL || (ix->table_val[s] >> 32) != 0) { bad = 1; continue; } ix->redge[4 * i + bb] = (u32)ix->table_val[s]; }
        if (l >> bb & 1) { u64 x = (first >> 2) | ((u64)bb << 58); u64 s = tab.find(to_device_form(x)); if (s == ~0ULL) { bad = 1; continue; } u32 nd = (u32)ix->table_val[s]; if ((u32)(ix->table_val[s] >> 32) != node_len[nd] - 1) { bad = 1; continue; } ix->ledge[4 * i + bb] = nd; }
      }
    }
  });
  phase("table values + edges");
  if (bad) { delete ix; return fail(NB_ERR_INVALID, "internal: de Bruijn edge does not land on a unitig boundary"); }
  *out = ix;
  return NB_OK;
}

extern "C" {

int nb_index_build_from_sequences(const uint8_t* seq_ascii, const uint64_t* seq_off, uint32_t n_seqs, int n_threads, nb_index** out) {
  if (!seq_off || !out || (!seq_ascii && n_seqs)) return fail(NB_ERR_INVALID, "null argument");
  std::vector<std::vector<u8>> seqs(n_seqs);
  for (u32 s = 0; s < n_seqs; s++) { u64 a = seq_off[s], b = seq_off[s + 1]; seqs[s].resize(b - a); for (u64 i = a; i < b; i++) seqs[s][i - a] = base_code(seq_ascii[i]); }
  return nb_build_index_impl(seqs, n_threads, out);
}
// utils::get_reference_sequence_data (src/utils.rs:7-24): DnaString::from_acgt_bytes over the sequence column, then build_index
int nb_index_build(const nb_library* lib, int n_threads, nb_index** out) {
  if (!lib || !out) return fail(NB_ERR_INVALID, "null argument");
  const std::vector<std::string>& col = lib->columns[lib->seq_idx];
  std::vector<std::vector<u8>> seqs(col.size());
  for (size_t s = 0; s < col.size(); s++) { seqs[s].resize(col[s].size()); for (size_t i = 0; i < col[s].size(); i++) seqs[s][i] = base_code((u8)col[s][i]); }
  return nb_build_index_impl(seqs, n_threads, out);
}
void nb_index_free(nb_index* ix) { delete ix; }
int nb_index_compare(const nb_index* a, const nb_index* b) {
  if (!a || !b) return fail(NB_ERR_INVALID, "null argument");
  if (a->n_kmers != b->n_kmers || a->unitig_bases != b->unitig_bases || a->n_sequences != b->n_sequences) return 1;
  if (a->unitig != b->unitig) return 2;
  if (a->node.size() != b->node.size() || (a->node.size() && memcmp(a->node.data(), b->node.data(), a->node.size() * sizeof(NodeRec)))) return 3;
  if (a->redge != b->redge || a->ledge != b->ledge) return 4;
  if (a->col_off != b->col_off || a->col_ids != b->col_ids) return 5;
  if (a->col_meta != b->col_meta) return 6;
  u64 na = 0, nbk = 0;
  const std::vector<u64>& bv = b->table_val;
  Table tb{b->table_key, nullptr, nullptr, b->table_buckets};
  for (u64 h = 0; h < b->table_key.size(); h++) nbk += b->table_key[h] >> 63;
  for (u64 h = 0; h < a->table_key.size(); h++) if (a->table_key[h] >> 63) {
    na++; u64 s = tb.find(a->table_key[h] & KMASK);
    if (s == ~0ULL || bv[s] != a->table_val[h]) return 7;
  }
  return (na == nbk && na == a->n_kmers) ? 0 : 7;
}
int nb_index_stats(const nb_index* ix, uint64_t* o) {
  if (!ix || !o) return fail(NB_ERR_INVALID, "null argument");
  o[0] = ix->n_kmers; o[1] = ix->node.size(); o[2] = ix->col_off.size() - 1; o[3] = ix->col_off.back(); o[4] = ix->unitig_bases;
  o[5] = ix->table_key.size(); o[6] = ix->device_bytes(); o[7] = ix->n_sequences;
  return NB_OK;
}
uint64_t nb_index_dump(const nb_index* ix, char* buf, uint64_t cap) {
  std::vector<std::string> lines(ix->node.size());
  for (size_t i = 0; i < ix->node.size(); i++) {
    const NodeRec& nr = ix->node[i]; u64 st = (u64)nr.start_lo | ((u64)(nr.exts_hi >> 8) << 32);
    std::string l(nr.len, 'A');
    for (u32 p = 0; p < nr.len; p++) { u64 pos = st + p; l[p] = "ACGT"[(ix->unitig[pos >> 5] >> (2 * (pos & 31))) & 3]; }
    l += "\t";
    for (u32 c = ix->col_off[nr.colour]; c < ix->col_off[nr.colour + 1]; c++) { if (c != ix->col_off[nr.colour]) l += ","; l += std::to_string(ix->col_ids[c]); }
    l += "\t" + std::to_string(nr.exts_hi & 15) + "\t" + std::to_string((nr.exts_hi >> 4) & 15) + "\n";
    lines[i] = l;
  }
  std::sort(lines.begin(), lines.end());
  u64 tot = 0; for (auto& l : lines) tot += l.size();
  if (buf && cap >= tot) { u64 p = 0; for (auto& l : lines) { memcpy(buf + p, l.data(), l.size()); p += l.size(); } }
  return tot;
}

}  // extern "C"

// ---- on-disk index cache: the flat arrays exactly as they are uploaded to HBM, followed by a 64-bit
// checksum of everything before it: a damaged cache would otherwise send the kernels out of bounds
namespace {
const char INDEX_MAGIC[8] = {'N', 'B', '2', 'I', 'D', 'X', '0', '6'};
struct Sum {
  u64 h = 0x243F6A8885A308D3ULL;
  void add(const void* p, size_t bytes) {
    const u8* b = (const u8*)p; size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { u64 w; memcpy(&w, b + i, 8); h = (h ^ w) * 0x9E3779B97F4A7C15ULL; h ^= h >> 29; }
    if (i < bytes) { u64 w = 0; memcpy(&w, b + i, bytes - i); h = (h ^ w) * 0x9E3779B97F4A7C15ULL; h ^= h >> 29; }
    h = (h ^ bytes) * 0xC2B2AE3D27D4EB4FULL; h ^= h >> 31;
  }
};
template <class T> bool put_vec(FILE* f, const std::vector<T>& v, Sum& sum) {
  u64 n = v.size(); sum.add(&n, 8); if (n) sum.add(v.data(), n * sizeof(T));
  return fwrite(&n, 8, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <class T> bool get_vec(FILE* f, std::vector<T>& v, Sum& sum, u64& left) {
  u64 n;
  if (left < 8 || fread(&n, 8, 1, f) != 1) return false;
  left -= 8;
  if (n > left / sizeof(T)) return false;   // an element count the file cannot hold (damaged header): refuse before allocating
  v.resize(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) return false;
  left -= n * sizeof(T);
  sum.add(&n, 8); if (n) sum.add(v.data(), n * sizeof(T));
  return true;
}
}
extern "C" int nb_index_save(const nb_index* ix, const char* path) {
  if (!ix || !path) return fail(NB_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "wb"); if (!f) return fail(NB_ERR_IO, std::string("could not open ") + path);
  u64 scalars[4] = {ix->table_buckets, ix->n_kmers, ix->unitig_bases, ix->n_sequences};
  Sum sum; sum.add(scalars, sizeof scalars);
  bool ok = fwrite(INDEX_MAGIC, 8, 1, f) == 1 && fwrite(scalars, 8, 4, f) == 4 && put_vec(f, ix->table_key, sum) && put_vec(f, ix->table_val, sum) && put_vec(f, ix->unitig, sum) &&
            put_vec(f, ix->node, sum) && put_vec(f, ix->redge, sum) && put_vec(f, ix->ledge, sum) && put_vec(f, ix->col_off, sum) && put_vec(f, ix->col_ids, sum) && put_vec(f, ix->col_meta, sum);
  ok = ok && fwrite(&sum.h, 8, 1, f) == 1;
  ok = (fclose(f) == 0) && ok;
  return ok ? NB_OK : fail(NB_ERR_IO, std::string("short write on ") + path);
}
extern "C" int nb_index_load(const char* path, nb_index** out) {
  if (!path || !out) return fail(NB_ERR_INVALID, "null argument");
  FILE* f = fopen(path, "rb"); if (!f) return fail(NB_ERR_IO, std::string("could not open ") + path);
  u64 left = 0;
  if (fseek(f, 0, SEEK_END) == 0) { long long sz = ftello(f); if (sz > 0) left = (u64)sz; rewind(f); }
  nb_index* ix = new nb_index(); char magic[8]; u64 scalars[4], want = 0; Sum sum;
  bool ok = left >= 48 && fread(magic, 8, 1, f) == 1 && !memcmp(magic, INDEX_MAGIC, 8) && fread(scalars, 8, 4, f) == 4;
  if (ok) { left -= 40; sum.add(scalars, sizeof scalars); }
  ok = ok && get_vec(f, ix->table_key, sum, left) && get_vec(f, ix->table_val, sum, left) && get_vec(f, ix->unitig, sum, left) &&
       get_vec(f, ix->node, sum, left) && get_vec(f, ix->redge, sum, left) && get_vec(f, ix->ledge, sum, left) && get_vec(f, ix->col_off, sum, left) && get_vec(f, ix->col_ids, sum, left) && get_vec(f, ix->col_meta, sum, left);
  ok = ok && left == 8 && fread(&want, 8, 1, f) == 1 && want == sum.h;
  fclose(f);
  if (ok) { ix->table_buckets = scalars[0]; ix->n_kmers = scalars[1]; ix->unitig_bases = scalars[2]; ix->n_sequences = scalars[3];
    ok = ix->table_buckets >= 1 && ix->table_key.size() == 4 * ix->table_buckets && ix->table_val.size() == ix->table_key.size() && ix->redge.size() == 4 * ix->node.size() && ix->ledge.size() == ix->redge.size() &&
         !ix->col_off.empty() && ix->col_meta.size() == 4 * (ix->col_off.size() - 1) && ix->unitig.size() >= (ix->unitig_bases + 31) / 32 + 2; }
  if (!ok) { delete ix; return fail(NB_ERR_PARSE, std::string("not a nimble_b200 index file (or truncated / damaged): ") + path); }
  *out = ix; return NB_OK;
}

// ---- index cache keyed by the library's sequences (the reference rebuilds its index on every run,
// src/reference_library.rs + debruijn_mapping build_index).  The index is a function of the sequence column alone, after
// DnaString::from_acgt_bytes' folding (case, non-ACGT -> A): the key hashes exactly that, two independent 64-bit sums, plus
// the artefact's format tag, so a cache written by another layout version is simply not found.
extern "C" int nb_index_cache_key(const nb_library* lib, char out_hex[33]) {
  if (!lib || !out_hex) return fail(NB_ERR_INVALID, "null argument");
  const std::vector<std::string>& col = lib->columns[lib->seq_idx];
  u64 h1 = 0x243F6A8885A308D3ULL, h2 = 0x13198A2E03707344ULL;
  auto mix = [&](u64 w) { h1 = (h1 ^ w) * 0x9E3779B97F4A7C15ULL; h1 ^= h1 >> 29; h2 = (h2 + w) * 0xC2B2AE3D27D4EB4FULL; h2 ^= h2 >> 32; };
  { u64 tag; memcpy(&tag, INDEX_MAGIC, 8); mix(tag); mix((u64)col.size()); }
  for (const std::string& q : col) {
    mix((u64)q.size());
    u64 w = 0; int n = 0;
    for (size_t i = 0; i < q.size(); i++) { w = (w << 2) | base_code((u8)q[i]); if (++n == 32) { mix(w); w = 0; n = 0; } }
    if (n) mix(w ^ ((u64)n << 58) ^ 0x8000000000000000ULL);
  }
  mix(h1 ^ (h2 >> 7));
  snprintf(out_hex, 33, "%016llx%016llx", (unsigned long long)h1, (unsigned long long)h2);
  return NB_OK;
}
// The index of `lib`: from <cache_dir>/<key>.nbix when that file is there and sound, else built on GPU `device` and written
// there (to a temporary name first: a reader never sees half a file).  cache_dir NULL: $NB_INDEX_CACHE; neither: just built.
// A cache that cannot be written is reported on stderr and otherwise ignored.
extern "C" int nb_index_build_cached(const nb_library* lib, const char* cache_dir, int device, int n_threads, nb_index** out) {
  if (!lib || !out) return fail(NB_ERR_INVALID, "null argument");
  if (!cache_dir || !*cache_dir) cache_dir = getenv("NB_INDEX_CACHE");
  if (!cache_dir || !*cache_dir) return nb_index_build_gpu(lib, device, n_threads, out);
  char key[33]; int rc = nb_index_cache_key(lib, key); if (rc) return rc;
  const std::string path = std::string(cache_dir) + "/" + key + ".nbix";
  if (nb_index_load(path.c_str(), out) == NB_OK) {
    if ((*out)->n_sequences == lib->columns[lib->seq_idx].size()) return NB_OK;
    nb_index_free(*out); *out = nullptr;              // (a key collision would have to get past this and the checksum)
  }
  rc = nb_index_build_gpu(lib, device, n_threads, out);
  if (rc) return rc;
  const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
  if (nb_index_save(*out, tmp.c_str()) != NB_OK || rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); fprintf(stderr, "nimble_b200: could not write the index cache %s (continuing without it)\n", path.c_str()); }
  return NB_OK;
}

