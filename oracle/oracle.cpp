// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY. NOT PART OF THE PRODUCT PATH.
//
// CPU restatement ("oracle") of the nimble-aligner read-alignment hot path. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// It shares no code with nimble_aligner_b200/ (own index builder, own walk, own string layer).
//
// Parity status: PINNED by the reference's own known-answer tests (tests/test_oracle_golden.py):
//   tests/basic-cases.rs:44-307 (8 goldens), tests/mismatch.rs:11-60 (2 goldens),
//   src/align.rs:1061-1107 (5 pseudoalign known answers), src/align.rs:1109-1752 (unit vectors for
//   filter_pair / roll-up / orientation filters / unmap / intersect / maxinfo / trim),
//   src/filter/align.rs:51-194, src/utils.rs:362-403 (entropy).
// UNPINNED (third-party semantics recalled from hextraza/rust-pseudoaligner + 10XGenomics/rust-debruijn,
// both un-vendored and un-versioned in Cargo.toml:22-23; no reference test exercises them):
//   seed stride 3, left-extension trigger floor(0.2*len) and its offset-0 quirk, cycle breaking in
//   unitig compaction (canonical rule here: a pure cycle starts at its smallest k-mer),
//   natural_lexical_cmp on non-ASCII names.
//
// What each block follows (all paths under /root/reference):
//   Index build      : debruijn_mapping::build_index::build_index::<Kmer30>  (call site src/bin/main.rs:121-128;
//                      semantics SURVEY.md Appendix A)
//   map_read         : Pseudoaligner::map_read_with_mismatch (call site src/align.rs:965; Appendix B)
//   shannon_entropy  : src/utils.rs:96-119
//   pseudoalign      : src/align.rs:945-989
//   filter_metrics   : src/filter/align.rs:4-45
//   maxinfo / trim   : src/align.rs:866-942
//   score_sequences  : src/align.rs:475-729
//   get_calls        : src/align.rs:392-467
//   orientation      : src/align.rs:144-375
//   roll-up / unmap  : src/align.rs:802-864
//   intersect        : src/align.rs:763-796 (+ array_tool 1.0.3 Intersect/Uniq semantics)
//   sort             : src/utils.rs:54-59, lexical-sort 0.3.1 natural_lexical_cmp (src/align.rs:846)
//   BAM scope rows   : src/process/bam.rs:305-405 (zero rows), 245-303
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

namespace {

const int K = 30;                      // src/align.rs:21 (Kmer30)
const u64 KMASK = (1ULL << 60) - 1;
const size_t MIN_READ_LENGTH = 40;     // src/align.rs:18
const double MIN_ENTROPY_SCORE = 1.75; // src/align.rs:19
const char* REV_SUFFIX = "\xC2\xA7rev"; // "§rev", src/reference_library.rs:8

// debruijn::dna_string base_to_bits: ACGT (either case) -> 0..3, anything else -> 0 ('A').
inline u8 base_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
  }
}
typedef std::vector<u8> Dna;
Dna dna_from_ascii(const char* s, size_t n) { Dna d(n); for (size_t i = 0; i < n; i++) d[i] = base_code(s[i]); return d; }
std::string dna_to_string(const Dna& d) { std::string s(d.size(), 'A'); for (size_t i = 0; i < d.size(); i++) s[i] = "ACGT"[d[i]]; return s; }
inline u64 kmer_at(const Dna& d, size_t pos) { u64 v = 0; for (int i = 0; i < K; i++) v = (v << 2) | d[pos + i]; return v; }

// ---------------------------------------------------------------- index (Appendix A)
struct Node {
  Dna seq;
  u32 colour;
  u8 lext, rext;          // bit b set <=> base b observed left of first k-mer / right of last k-mer
  int32_t redge[4], ledge[4];
};
// Open-addressed k-mer map (test infrastructure of its own: the 200k-transcript parity library has 2e8 k-mers, far
// beyond what node-based std::unordered_map holds in reasonable time / memory).  Keys are 60-bit k-mers; ~0 = empty.
struct KmerMap {
  struct E { u64 key, val; };
  std::vector<E> e; u64 mask = 0, n = 0;
  void init(size_t expect) { size_t c = 16; while (c < expect * 2) c <<= 1; e.assign(c, E{~0ULL, 0}); mask = c - 1; n = 0; }
  static u64 h(u64 k) { k ^= k >> 31; k *= 0x9E3779B97F4A7C15ULL; k ^= k >> 29; return k; }
  void put(u64 k, u64 v) { u64 i = h(k) & mask; while (e[i].key != ~0ULL && e[i].key != k) i = (i + 1) & mask; if (e[i].key == ~0ULL) n++; e[i].key = k; e[i].val = v; }
  bool get(u64 k, u64& v) const { if (e.empty()) return false; u64 i = h(k) & mask; while (e[i].key != ~0ULL) { if (e[i].key == k) { v = e[i].val; return true; } i = (i + 1) & mask; } return false; }
  u64 at(u64 k) const { u64 v = 0; if (!get(k, v)) abort(); return v; }
};
struct Index {
  std::vector<Node> nodes;
  std::vector<std::vector<u32>> colours;                   // eq_classes
  KmerMap kmap;                                            // k-mer -> node | offset << 32
  u64 n_kmers = 0;
};

struct Occ { u64 kmer; u32 id; u8 l, r; };

void build_index(const std::vector<Dna>& seqs, Index& ix) {
  const bool tm = getenv("ORC_TIMING") != nullptr; auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) { if (!tm) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "oracle build_index: %s %.2fs\n", what, std::chrono::duration<double>(t - t_last).count()); t_last = t; };
  std::vector<Occ> occ;
  { size_t tot = 0; for (auto& d : seqs) if (d.size() >= (size_t)K) tot += d.size() - K + 1; occ.reserve(tot); }
  for (size_t s = 0; s < seqs.size(); s++) {
    const Dna& d = seqs[s];
    if (d.size() < (size_t)K) continue;
    u64 km = 0;
    for (size_t p = 0; p < d.size(); p++) {
      km = ((km << 2) | d[p]) & KMASK;
      if (p + 1 < (size_t)K) continue;
      size_t st = p + 1 - K;
      Occ o; o.kmer = km; o.id = (u32)s;
      o.l = st > 0 ? d[st - 1] : 4; o.r = p + 1 < d.size() ? d[p + 1] : 4;
      occ.push_back(o);
    }
  }
  lap("enumerate");
  auto occ_less = [](const Occ& a, const Occ& b) { return a.kmer != b.kmer ? a.kmer < b.kmer : a.id < b.id; };
  {  // sort by (k-mer, id); large inputs: partition by the k-mer's top byte, sort the parts on threads
    unsigned T = std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
    if (occ.size() < (1u << 20) || T == 1) std::sort(occ.begin(), occ.end(), occ_less);
    else {
      std::vector<size_t> cnt(257, 0);
      for (auto& o : occ) cnt[(o.kmer >> 52) + 1]++;
      for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
      std::vector<Occ> tmp(occ.size()); std::vector<size_t> cur(cnt.begin(), cnt.end() - 1);
      for (auto& o : occ) tmp[cur[o.kmer >> 52]++] = o;
      occ.swap(tmp); tmp.clear(); tmp.shrink_to_fit();
      std::atomic<int> next(0); std::vector<std::thread> th;
      for (unsigned t = 0; t < T; t++) th.emplace_back([&]() { for (int b; (b = next++) < 256;) std::sort(occ.begin() + cnt[b], occ.begin() + cnt[b + 1], occ_less); });
      for (auto& t : th) t.join();
    }
  }
  lap("sort");
  std::vector<u64> kmers; std::vector<u8> L, R; std::vector<u32> col;
  // colour interning: open-addressed table over a hash of the id list, every hit verified element-wise
  kmers.reserve(occ.size() / 2); L.reserve(occ.size() / 2); R.reserve(occ.size() / 2); col.reserve(occ.size() / 2);
  std::vector<u32> itab(1u << 16, 0xFFFFFFFFu); u64 imask = itab.size() - 1;
  auto id_hash = [](const u32* p, size_t n) { u64 hv = 0x243F6A8885A308D3ULL ^ n; for (size_t i = 0; i < n; i++) { hv = (hv ^ p[i]) * 0x9E3779B97F4A7C15ULL; hv ^= hv >> 32; } return hv; };
  std::vector<u32> ids;
  for (size_t i = 0; i < occ.size();) {
    size_t j = i; u8 l = 0, r = 0; ids.clear();
    while (j < occ.size() && occ[j].kmer == occ[i].kmer) {
      if (occ[j].l < 4) l |= 1 << occ[j].l;
      if (occ[j].r < 4) r |= 1 << occ[j].r;
      if (ids.empty() || ids.back() != occ[j].id) ids.push_back(occ[j].id);
      j++;
    }
    u64 h = id_hash(ids.data(), ids.size()) & imask;
    u32 c = 0xFFFFFFFFu;
    for (;; h = (h + 1) & imask) {
      u32 q = itab[h];
      if (q == 0xFFFFFFFFu) break;
      const std::vector<u32>& cq = ix.colours[q];
      if (cq.size() == ids.size() && std::equal(cq.begin(), cq.end(), ids.begin())) { c = q; break; }
    }
    if (c == 0xFFFFFFFFu) {
      c = (u32)ix.colours.size(); itab[h] = c; ix.colours.push_back(ids);
      if (ix.colours.size() * 2 > itab.size()) {   // grow + rehash
        itab.assign(itab.size() * 4, 0xFFFFFFFFu); imask = itab.size() - 1;
        for (u32 q = 0; q < ix.colours.size(); q++) { u64 g = id_hash(ix.colours[q].data(), ix.colours[q].size()) & imask; while (itab[g] != 0xFFFFFFFFu) g = (g + 1) & imask; itab[g] = q; }
      }
    }
    kmers.push_back(occ[i].kmer); L.push_back(l); R.push_back(r); col.push_back(c);
    i = j;
  }
  occ.clear(); occ.shrink_to_fit();
  lap("group + colours");
  size_t n = kmers.size();
  ix.n_kmers = n;
  KmerMap kidx; kidx.init(n);
  for (size_t i = 0; i < n; i++) kidx.put(kmers[i], (u64)i);
  lap("k-mer index");
  std::vector<int32_t> succ(n, -1), pred(n, -1);
  for (size_t i = 0; i < n; i++) {
    if (__builtin_popcount(R[i]) != 1) continue;
    int b = __builtin_ctz(R[i]);
    u64 y = ((kmers[i] << 2) | (u64)b) & KMASK;
    u32 j = (u32)kidx.at(y);
    if (__builtin_popcount(L[j]) == 1 && col[i] == col[j]) { succ[i] = (int32_t)j; pred[j] = (int32_t)i; }
  }
  lap("succ/pred");
  std::vector<int32_t> start_node(n, -1), end_node(n, -1);
  std::vector<char> used(n, 0);
  ix.kmap.init(n);
  auto emit = [&](size_t s) {
    Node nd; nd.colour = col[s]; nd.lext = L[s];
    for (int i = K - 1; i >= 0; i--) nd.seq.push_back((u8)((kmers[s] >> (2 * i)) & 3));
    size_t cur = s; used[cur] = 1;
    u32 id = (u32)ix.nodes.size();
    u32 off = 0;
    ix.kmap.put(kmers[cur], (u64)id | ((u64)off << 32));
    while (succ[cur] >= 0 && !used[succ[cur]]) {
      cur = (size_t)succ[cur]; used[cur] = 1; off++;
      nd.seq.push_back((u8)(kmers[cur] & 3));
      ix.kmap.put(kmers[cur], (u64)id | ((u64)off << 32));
    }
    nd.rext = R[cur];
    for (int b = 0; b < 4; b++) nd.redge[b] = nd.ledge[b] = -1;
    start_node[s] = (int32_t)id; end_node[cur] = (int32_t)id;
    ix.nodes.push_back(std::move(nd));
  };
  for (size_t i = 0; i < n; i++) if (pred[i] < 0) emit(i);
  for (size_t i = 0; i < n; i++) if (!used[i]) emit(i);   // pure cycles: smallest k-mer first (arrays are k-mer sorted)
  lap("unitigs");
  for (size_t i = 0; i < n; i++) {
    if (start_node[i] >= 0) {
      Node& nd = ix.nodes[start_node[i]];
      for (int b = 0; b < 4; b++) if (nd.lext >> b & 1) {
        u64 x = (kmers[i] >> 2) | ((u64)b << 58);
        nd.ledge[b] = end_node[kidx.at(x)];
      }
    }
    if (end_node[i] >= 0) {
      Node& nd = ix.nodes[end_node[i]];
      for (int b = 0; b < 4; b++) if (nd.rext >> b & 1) {
        u64 y = ((kmers[i] << 2) | (u64)b) & KMASK;
        nd.redge[b] = start_node[kidx.at(y)];
      }
    }
  }
}

struct Work { u64 probes = 0, nodes = 0, bases = 0, colour_elems = 0, reads = 0, in_bases = 0; };

// Appendix B. Returns false for None.
bool map_read(const Index& ix, const Dna& read, size_t allowed, std::vector<u32>& eq, size_t& cov, size_t& mm, Work& w) {
  size_t n = read.size();
  cov = 0; mm = 0; eq.clear();
  if (n < (size_t)K) return false;
  size_t last_kpos = n - K;
  std::vector<u32> visited;
  size_t left_thresh = (size_t)(0.2 * (double)n);
  size_t kp = 0;
  bool found = false; u32 node = 0, off = 0;
  auto find = [&](size_t& pos) -> bool {
    while (pos <= last_kpos) {
      w.probes++;
      u64 v;
      if (ix.kmap.get(kmer_at(read, pos), v)) { node = (u32)v; off = (u32)(v >> 32); return true; }
      pos += 3;
    }
    return false;
  };
  found = find(kp);
  if (found && kp >= left_thresh) {
    size_t lp = kp - 1; u32 pn = node; size_t po = off > 0 ? off - 1 : 0;
    for (;;) {
      const Node& nd = ix.nodes[pn];
      size_t m = std::min(lp + 1, po + 1), snp = 0, mb = 0; bool brk = false;
      for (size_t i = 0; i < m; i++) {
        w.bases++;
        if (nd.seq[po - i] != read[lp - i]) { mm++; snp++; if (snp > allowed) { brk = true; break; } }
        mb++; cov++;
      }
      if (lp + 1 - mb == 0 || brk) break;
      lp -= mb;
      u8 b = read[lp];
      if (nd.lext >> b & 1) {
        pn = (u32)nd.ledge[b]; po = ix.nodes[pn].seq.size() - K; visited.push_back(pn); w.nodes++;
      } else break;
    }
  }
  if (found && kp <= last_kpos) {
    for (;;) {
      const Node& nd = ix.nodes[node];
      kp += K; cov += K; visited.push_back(node); w.nodes++;
      size_t ro = off + K, m = std::min(n - kp, nd.seq.size() - ro), snp = 0, mb = 0; bool brk = false;
      for (size_t i = 0; i < m; i++) {
        w.bases++;
        if (nd.seq[ro + i] != read[kp + i]) { mm++; snp++; if (snp > allowed) { brk = true; break; } }
        mb++; cov++;
      }
      kp += mb;
      if (kp >= n) break;
      u8 b = read[kp];
      if (!brk && (nd.rext >> b & 1)) { node = (u32)nd.redge[b]; off = 0; kp -= K - 1; cov -= K - 1; }
      else { if (kp > last_kpos) break; if (!find(kp)) break; }
    }
  }
  if (visited.empty()) return false;
  // nodes_to_eq_class: sort by colour size, merge-intersect
  std::stable_sort(visited.begin(), visited.end(), [&](u32 a, u32 b) { return ix.colours[ix.nodes[a].colour].size() < ix.colours[ix.nodes[b].colour].size(); });
  eq = ix.colours[ix.nodes[visited[0]].colour]; w.colour_elems += eq.size();
  for (size_t i = 1; i < visited.size(); i++) {
    const std::vector<u32>& c = ix.colours[ix.nodes[visited[i]].colour]; w.colour_elems += c.size();
    std::vector<u32> out; std::set_intersection(eq.begin(), eq.end(), c.begin(), c.end(), std::back_inserter(out)); eq.swap(out);
  }
  return true;
}

// ---------------------------------------------------------------- nimble layer
enum Reason : u8 {  // src/align.rs:32-51, same order
  ScoreBelowThreshold, DiscardedMultipleMatch, DiscardedNonzeroMismatch, NoMatch, NoMatchAndScoreBelowThreshold,
  DifferentFilterReasons, NotMatchingPair, ForceIntersectFailure, ShortRead, MaxHitsExceeded, HighEntropy,
  SuccessfulMatch, StrandWasWrong, TriageEmptyEquivalenceClass, AboveMismatchThreshold, SkippedAlignDueToUnpairedDummy, ReasonNone
};
enum Chem { Unstranded = 0, FivePrime = 1, ThreePrime = 2, ChemNone = 3 };

struct Cfg {  // AlignFilterConfig, src/align.rs:80-95 (POD mirror; must match oracle/__init__.py OrcCfg)
  double score_percent; u64 score_threshold; u64 num_mismatches; int32_t discard_multiple_matches; int32_t require_valid_pair;
  u64 discard_multi_hits; u64 max_hits_to_report; int32_t intersect_level; int32_t strand_filter;
  u64 trim_target_length; double trim_strictness; int32_t group_header_is_nt_sequence; int32_t faithful_cost;
};

struct Lib {  // Reference, src/reference_library.rs:11-17 — only the three columns the hot path reads
  std::vector<std::string> names, groups; std::vector<Dna> seqs;
};

double shannon_entropy(const std::string& dna) {  // src/utils.rs:96-119
  double total = (double)dna.size(); double f[4] = {0, 0, 0, 0};
  for (char c : dna) switch (c) { case 'A': f[0] += 1; break; case 'T': f[1] += 1; break; case 'C': f[2] += 1; break; case 'G': f[3] += 1; break; default: break; }
  for (int i = 0; i < 4; i++) f[i] /= total;
  double e = 0.0;
  for (int i = 0; i < 4; i++) if (f[i] > 0.0) e += f[i] * std::log2(f[i]);
  return -e;
}

struct MaxinfoTables { std::vector<i64> ls, qp; };
i64 f64_to_i64_sat(double v) {  // Rust `as i64`: saturating, NaN -> 0
  if (std::isnan(v)) return 0;
  if (v >= 9223372036854775807.0) return std::numeric_limits<i64>::max();
  if (v <= -9223372036854775808.0) return std::numeric_limits<i64>::min();
  return (i64)v;
}
void maxinfo_tables(size_t target, double strictness, MaxinfoTables& t) {  // src/align.rs:873-897
  const size_t LONGEST = 1000, MAXQ = 60;
  std::vector<double> ls(LONGEST), qp(MAXQ + 1);
  for (size_t i = 0; i < LONGEST; i++) {
    double pow1 = std::exp((double)target - (double)i - 1.0);
    double unique = std::log(1.0 / (1.0 + pow1));
    double coverage = std::log((double)(i + 1)) * (1.0 - strictness);
    ls[i] = unique + coverage;
  }
  for (size_t i = 0; i <= MAXQ; i++) { double pc = 1.0 - std::pow(10.0, -((0.5 + (double)i) / 10.0)); qp[i] = std::log(pc) * strictness; }
  auto norm_ratio = [](const std::vector<double>& a, size_t margin) {
    double mx = std::fabs(a[0]); for (size_t i = 1; i < a.size(); i++) { double v = std::fabs(a[i]); if (v > mx) mx = v; }
    return 9223372036854775807.0 / (mx * (double)margin);
  };
  double r1 = norm_ratio(ls, LONGEST * 2), r2 = norm_ratio(qp, LONGEST * 2);
  double ratio = std::fmax(r1, r2);  // f64::max ignores NaN like fmax
  t.ls.resize(LONGEST); t.qp.resize(MAXQ + 1);
  for (size_t i = 0; i < LONGEST; i++) t.ls[i] = f64_to_i64_sat(ls[i] * ratio);
  for (size_t i = 0; i <= MAXQ; i++) t.qp[i] = f64_to_i64_sat(qp[i] * ratio);
}
size_t maxinfo_scan(const u8* q, size_t n, const MaxinfoTables& t) {  // src/align.rs:899-924
  i64 acc = 0; double max_score = -std::numeric_limits<double>::max(); size_t pos = 0;
  for (size_t i = 0; i < n; i++) {
    size_t qq = q[i]; if (qq > 60) qq = 60;
    acc = (i64)((u64)acc + (u64)t.qp[qq]);
    i64 ls = i < t.ls.size() ? t.ls[i] : 0;
    i64 score = (i64)((u64)ls + (u64)acc);
    if ((double)score >= max_score) { max_score = (double)score; pos = i + 1; }
  }
  if (pos < 1 || max_score == 0.0) return 0;
  return pos < n ? pos : n;
}
size_t maxinfo(const u8* q, size_t n, size_t target, double strictness) {  // faithful: rebuilds the tables per call
  MaxinfoTables t; maxinfo_tables(target, strictness, t); return maxinfo_scan(q, n, t);
}

struct ReadRec { u8 reason; u32 score; u32 mm; u32 trimmed_len; std::vector<u32> ec; bool pass; double norm; };

// src/align.rs:945-989 + src/filter/align.rs:4-45. rec.ec = raw eq-class from the walk (even when filtered).
void pseudoalign(const Index& ix, const Dna& seq, const Cfg& cfg, size_t min_len, ReadRec& rec, Work& w) {
  rec.reason = SuccessfulMatch; rec.score = 0; rec.mm = 0; rec.ec.clear(); rec.pass = false; rec.norm = 0.0;
  rec.trimmed_len = (u32)seq.size();
  if (seq.size() < min_len) { rec.reason = ShortRead; return; }
  if (shannon_entropy(dna_to_string(seq)) < MIN_ENTROPY_SCORE) { rec.reason = HighEntropy; return; }
  size_t cov, mm;
  w.reads++; w.in_bases += seq.size();
  if (!map_read(ix, seq, (size_t)cfg.num_mismatches, rec.ec, cov, mm, w)) { rec.reason = NoMatch; return; }
  rec.score = (u32)cov; rec.mm = (u32)mm;
  double norm = (double)cov / (double)seq.size(); rec.norm = norm;
  if (cov >= cfg.score_threshold && norm >= cfg.score_percent && !rec.ec.empty()) {
    if (cfg.discard_multiple_matches && rec.ec.size() > 1) rec.reason = DiscardedMultipleMatch;
    else if (mm > cfg.num_mismatches) rec.reason = AboveMismatchThreshold;
    else { rec.reason = SuccessfulMatch; rec.pass = true; }
  } else rec.reason = ScoreBelowThreshold;
}

// lexical-sort 0.3.1 natural_lexical_cmp restated for ASCII: case-folded compare with digit runs compared by value;
// ties fall back to byte order so distinct strings never compare equal.
int natural_lexical_cmp(const std::string& a, const std::string& b) {
  size_t i = 0, j = 0;
  auto fold = [](unsigned char c) -> unsigned char { return (c >= 'A' && c <= 'Z') ? (unsigned char)(c + 32) : c; };
  while (i < a.size() && j < b.size()) {
    unsigned char ca = (unsigned char)a[i], cb = (unsigned char)b[j];
    if (ca >= '0' && ca <= '9' && cb >= '0' && cb <= '9') {
      size_t i0 = i, j0 = j;
      while (i0 < a.size() && a[i0] == '0') i0++;
      while (j0 < b.size() && b[j0] == '0') j0++;
      size_t i1 = i0, j1 = j0;
      while (i1 < a.size() && a[i1] >= '0' && a[i1] <= '9') i1++;
      while (j1 < b.size() && b[j1] >= '0' && b[j1] <= '9') j1++;
      // leading zeros consumed but a run of only zeros keeps i0 at the first non-zero char (value 0)
      size_t la = i1 - i0, lb = j1 - j0;
      if (la != lb) return la < lb ? -1 : 1;
      int c = a.compare(i0, la, b, j0, lb);
      if (c != 0) return c < 0 ? -1 : 1;
      i = i1; j = j1;
    } else {
      unsigned char fa = fold(ca), fb = fold(cb);
      if (fa != fb) return fa < fb ? -1 : 1;
      i++; j++;
    }
  }
  if (i < a.size()) return 1;
  if (j < b.size()) return -1;
  int c = a.compare(b);
  return c < 0 ? -1 : (c > 0 ? 1 : 0);
}

typedef std::vector<std::string> Strs;

bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

// src/align.rs:802-849
Strs feature_list(const std::vector<u32>& ec, const Lib& lib, const Cfg& cfg, bool ignore_rollup) {
  Strs res;
  if (ignore_rollup || cfg.group_header_is_nt_sequence) {
    for (u32 r : ec) res.push_back(lib.names[r]);
  } else {
    for (u32 r : ec) {
      const std::string* g = &lib.groups[r];
      if (g->empty()) g = &lib.names[r];
      if (std::find(res.begin(), res.end(), *g) == res.end()) res.push_back(*g);
    }
  }
  if (!ignore_rollup && cfg.discard_multi_hits > 0 && res.size() > cfg.discard_multi_hits) return Strs();
  std::sort(res.begin(), res.end(), [](const std::string& a, const std::string& b) { return natural_lexical_cmp(a, b) < 0; });
  return res;
}

// src/align.rs:851-864. Returns false where the reference panics ("Feature not found in reference columns").
bool unmap(const Strs& feats, const Lib& lib, const std::unordered_map<std::string, u32>* first_row, std::vector<u32>& out) {
  out.clear();
  for (const std::string& f : feats) {
    if (first_row) {  // same answer as the linear scan, without its cost
      auto it = first_row->find(f); if (it == first_row->end()) return false; out.push_back(it->second);
    } else {
      size_t p = 0; for (; p < lib.names.size(); p++) if (lib.names[p] == f) break;
      if (p == lib.names.size()) return false;
      out.push_back((u32)p);
    }
  }
  return true;
}

// src/align.rs:144-171
Strs filter_read_calls_with_orientation(const Strs& cls) {
  std::unordered_set<std::string> seen, to_remove;
  auto base_of = [](const std::string& f) { return ends_with(f, REV_SUFFIX) ? f.substr(0, f.size() - strlen(REV_SUFFIX)) : f; };
  for (const std::string& f : cls) { std::string b = base_of(f); if (seen.count(b)) to_remove.insert(b); else seen.insert(b); }
  Strs out;
  for (const std::string& f : cls) if (!to_remove.count(base_of(f))) out.push_back(f);
  return out;
}
typedef std::pair<std::string, bool> Call;
// src/align.rs:276-285
std::vector<Call> parse_calls(const Strs& calls) {
  std::vector<Call> out;
  for (const std::string& c : calls) {
    if (ends_with(c, "rev")) {
      std::string b = c;
      while (ends_with(b, "rev")) b.resize(b.size() - 3);
      while (ends_with(b, "\xC2\xA7")) b.resize(b.size() - 2);
      out.push_back({b, true});
    } else out.push_back({c, false});
  }
  return out;
}
// src/align.rs:287-309
void filter_unstranded(const std::vector<Call>& a, const std::vector<Call>& b, std::vector<Call>& fa, std::vector<Call>& fb) {
  fa.clear(); fb.clear();
  for (const Call& c : a) if (std::find(b.begin(), b.end(), c) == b.end()) fa.push_back(c);
  for (const Call& c : b) if (std::find(a.begin(), a.end(), c) == a.end()) fb.push_back(c);
}
// src/align.rs:311-375; five==true -> filter_five_prime, else filter_three_prime
void filter_prime(const std::vector<Call>& a, const std::vector<Call>& b, bool five, Strs& oa, Strs& ob) {
  std::vector<Call> ua, ub; filter_unstranded(a, b, ua, ub);
  std::vector<Call> seq_f, mate_f = ub;
  for (const Call& c : ua) {
    bool drop = five ? c.second : !c.second;
    if (drop) {
      for (size_t p = 0; p < mate_f.size(); p++) if (mate_f[p].first == c.first) { mate_f.erase(mate_f.begin() + p); break; }
    } else seq_f.push_back(c);
  }
  std::vector<Call> kept;
  for (const Call& m : mate_f) {
    bool need = five ? !m.second : m.second;
    if (need) { bool any = false; for (const Call& s : seq_f) if (s.first == m.first) { any = true; break; } if (any) kept.push_back(m); }
    else kept.push_back(m);
  }
  oa.clear(); ob.clear();
  for (const Call& c : seq_f) oa.push_back(c.first);
  for (const Call& c : kept) ob.push_back(c.first);
}
// array_tool 1.0.3 Intersect: unique(A) kept when present in B, in A's order
Strs at_intersect(const Strs& a, const Strs& b) {
  Strs ua; for (const std::string& x : a) if (std::find(ua.begin(), ua.end(), x) == ua.end()) ua.push_back(x);
  Strs out; for (const std::string& x : ua) if (std::find(b.begin(), b.end(), x) != b.end()) out.push_back(x);
  return out;
}

struct PairEntry { std::vector<u32> ec1, ec2; bool has1 = false, has2 = false; u64 pair_index = 0; };

// src/align.rs:178-252. Returns triage reason (ReasonNone when the callset was counted); callset in `out`.
// error=true where the reference would panic in unmap.
u8 coerce(const PairEntry& e, const Lib& lib, const Cfg& cfg, const std::unordered_map<std::string, u32>* first_row, Strs& out, bool& error) {
  Strs sf, mf;
  if (e.has1) sf = feature_list(e.ec1, lib, cfg, true);
  if (e.has2) mf = feature_list(e.ec2, lib, cfg, true);
  sf = filter_read_calls_with_orientation(sf);
  mf = filter_read_calls_with_orientation(mf);
  std::vector<Call> ps = parse_calls(sf), pm = parse_calls(mf);
  Strs a, b;
  switch (cfg.strand_filter) {
    case ChemNone: for (auto& c : ps) a.push_back(c.first); for (auto& c : pm) b.push_back(c.first); break;
    case Unstranded: { std::vector<Call> fa, fb; filter_unstranded(ps, pm, fa, fb); for (auto& c : fa) a.push_back(c.first); for (auto& c : fb) b.push_back(c.first); break; }
    case FivePrime: filter_prime(ps, pm, true, a, b); break;
    default: filter_prime(ps, pm, false, a, b); break;
  }
  Strs fin;
  if (cfg.intersect_level == 0) { fin = a; fin.insert(fin.end(), b.begin(), b.end()); }  // unique() result discarded, src/align.rs:794
  else {
    fin = at_intersect(a, b);
    if (fin.empty() && cfg.intersect_level == 1) { fin = a; fin.insert(fin.end(), b.begin(), b.end()); }
    // level 2 failure: ForceIntersectFailure is recorded, then overwritten by TriageEmptyEquivalenceClass below (src/align.rs:239-241)
  }
  std::vector<u32> rows;
  if (!unmap(fin, lib, first_row, rows)) { error = true; return ReasonNone; }
  out = feature_list(rows, lib, cfg, false);
  if (out.size() > cfg.max_hits_to_report) return MaxHitsExceeded;
  if (out.empty()) return TriageEmptyEquivalenceClass;
  return ReasonNone;
}

// src/align.rs:732-760
bool filter_pair(const std::vector<u32>& a0, const std::vector<u32>& b0) {
  if (a0.empty() || b0.empty()) return true;
  std::vector<u32> a = a0, b = b0; std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
  size_t matching = 0; for (size_t i = 0; i < std::min(a.size(), b.size()); i++) if (a[i] == b[i]) matching++;
  return matching != a.size() || matching != b.size();
}

struct PairRec { u8 triage; u8 fr1, fr2; u8 counted; u32 score1, score2; int32_t callset; };  // callset: index into the scope's result list, -1 none

struct ScopeResult {
  std::vector<std::pair<Strs, i64>> counts;  // sorted by Vec<String> Ord
  u64 n_pairs = 0;
};

struct Oracle {
  Lib lib; Cfg cfg; Index ix; std::unordered_map<std::string, u32> first_row; bool error = false;
  // last run outputs
  std::vector<ReadRec> reads;       // 2 per pair when mates present (seq, mate), else 1
  std::vector<PairRec> pairs;
  std::vector<ScopeResult> scopes;
  Work work;
  std::string blob;                 // serialised results
};

struct Input {
  const char* r1; const u64* r1_off; const char* r2; const u64* r2_off;   // ascii + offsets (n+1); r2 null => single-end
  const u8* q1; const u8* q2;                                            // raw phred bytes with the same offsets; null => no trimming (FASTQ mode)
  const u8* skip1; const u8* skip2;                                      // SKIP_ALIGN flags per pair side; null => none
};

// get_calls over pairs [p0,p1) as one aggregation scope (src/align.rs:392-467 + 475-729).
void run_scope(Oracle& o, const Input& in, u64 p0, u64 p1, ScopeResult& res, Work& w, const MaxinfoTables* tables) {
  const Cfg& cfg = o.cfg; bool paired = in.r2 != nullptr;
  std::unordered_map<std::string, PairEntry> score_map;
  std::unordered_map<std::string, std::pair<u8, u8>> filter_reasons;
  std::vector<std::string> keys(p1 - p0);
  for (u64 p = p0; p < p1; p++) {
    Dna read = dna_from_ascii(in.r1 + in.r1_off[p], in.r1_off[p + 1] - in.r1_off[p]);
    auto trimmed = [&](const Dna& d, const u8* q, u64 qoff) -> Dna {
      if (!q) return d;
      size_t tl = tables ? maxinfo_scan(q + qoff, d.size(), *tables) : maxinfo(q + qoff, d.size(), (size_t)cfg.trim_target_length, cfg.trim_strictness);
      return Dna(d.begin(), d.begin() + tl);
    };
    ReadRec& r1 = o.reads[paired ? 2 * p : p];
    if (in.skip1 && in.skip1[p]) { r1 = ReadRec(); r1.reason = SkippedAlignDueToUnpairedDummy; r1.score = r1.mm = 0; r1.pass = false; r1.trimmed_len = (u32)read.size(); }
    else pseudoalign(o.ix, trimmed(read, in.q1, in.r1_off[p]), cfg, MIN_READ_LENGTH, r1, w);
    std::string key = dna_to_string(read);
    std::vector<u32> ec1 = r1.pass ? r1.ec : std::vector<u32>(), ec2;
    u32 s1 = r1.pass ? r1.score : 0, s2 = 0;
    ReadRec* r2 = nullptr;
    if (paired) {
      Dna mate = dna_from_ascii(in.r2 + in.r2_off[p], in.r2_off[p + 1] - in.r2_off[p]);
      r2 = &o.reads[2 * p + 1];
      if (in.skip2 && in.skip2[p]) { *r2 = ReadRec(); r2->reason = SkippedAlignDueToUnpairedDummy; r2->score = r2->mm = 0; r2->pass = false; r2->trimmed_len = (u32)mate.size(); }
      else pseudoalign(o.ix, trimmed(mate, in.q2, in.r2_off[p]), cfg, MIN_READ_LENGTH, *r2, w);
      if (r2->pass) { ec2 = r2->ec; s2 = r2->score; }
      key += dna_to_string(mate);
    }
    PairRec& pr = o.pairs[p]; pr.triage = ReasonNone; pr.counted = 0; pr.callset = -1; pr.score1 = s1; pr.score2 = s2;
    keys[p - p0] = key;
    if (paired && cfg.require_valid_pair && filter_pair(ec1, ec2)) {
      pr.fr1 = pr.fr2 = NotMatchingPair; filter_reasons[key] = {NotMatchingPair, NotMatchingPair};
      continue;
    }
    pr.fr1 = r1.reason; pr.fr2 = paired ? r2->reason : (u8)SuccessfulMatch;  // mate_sequence_filter_reason None -> SuccessfulMatch (src/align.rs:596-599)
    filter_reasons[key] = {pr.fr1, pr.fr2};
    if (!ec1.empty() || !ec2.empty()) {
      PairEntry e; e.has1 = !ec1.empty(); e.has2 = !ec2.empty(); e.ec1 = ec1; e.ec2 = ec2; e.pair_index = p;
      score_map[key] = std::move(e);   // later duplicates overwrite (src/align.rs:685)
    }
  }
  // one vote per unique read_key (src/align.rs:440-449)
  std::map<Strs, i64> results;
  std::unordered_map<std::string, std::pair<u8, Strs>> per_key;
  for (auto& kv : score_map) {
    Strs cs; bool err = false;
    u8 tri = coerce(kv.second, o.lib, cfg, cfg.faithful_cost ? nullptr : &o.first_row, cs, err);
    if (err) { o.error = true; continue; }
    if (tri == ReasonNone) results[cs] += 1;
    per_key[kv.first] = {tri, cs};
  }
  res.counts.assign(results.begin(), results.end());   // std::map<vector<string>> order == Vec<String> Ord (bytewise)
  res.n_pairs = p1 - p0;
  // per-pair projection: every pair sharing a key reports that key's triage / callset
  for (u64 p = p0; p < p1; p++) {
    PairRec& pr = o.pairs[p];
    // filter_reasons is keyed by read_key and overwritten by later pairs (src/align.rs:586-600); the BAM driver looks
    // the reasons up by key (src/process/bam.rs:356-361), so every pair reports its key's last entry
    auto fr = filter_reasons.find(keys[p - p0]);
    if (fr != filter_reasons.end()) { pr.fr1 = fr->second.first; pr.fr2 = fr->second.second; }
    auto it = per_key.find(keys[p - p0]);
    if (it == per_key.end()) continue;
    pr.triage = it->second.first;
    if (pr.triage == ReasonNone) {
      pr.counted = 1;
      auto pos = std::lower_bound(res.counts.begin(), res.counts.end(), it->second.second, [](const std::pair<Strs, i64>& a, const Strs& b) { return a.first < b; });
      pr.callset = (int32_t)(pos - res.counts.begin());
    }
  }
}

void append_u64(std::string& s, u64 v) { s.append((const char*)&v, 8); }

}  // namespace

extern "C" {

// names/groups/seqs blobs: n_rows NUL-terminated strings each (rows already include the §rev rows, built by the caller
// the way src/reference_library.rs:130-153 does).
void* orc_create(u32 n_rows, const char* names, const char* groups, const char* seqs, const Cfg* cfg) {
  Oracle* o = new Oracle(); o->cfg = *cfg;
  const char* pn = names; const char* pg = groups; const char* ps = seqs;
  for (u32 i = 0; i < n_rows; i++) {
    o->lib.names.emplace_back(pn); pn += o->lib.names.back().size() + 1;
    o->lib.groups.emplace_back(pg); pg += o->lib.groups.back().size() + 1;
    size_t n = strlen(ps); o->lib.seqs.push_back(dna_from_ascii(ps, n)); ps += n + 1;   // utils.rs:7-24 from_acgt_bytes
    o->first_row.emplace(o->lib.names.back(), i);   // emplace keeps the first row of a duplicated name (== position())
  }
  build_index(o->lib.seqs, o->ix);
  return o;
}
void orc_free(void* h) { delete (Oracle*)h; }
void orc_set_cfg(void* h, const Cfg* cfg) { ((Oracle*)h)->cfg = *cfg; }

void orc_index_stats(void* h, u64* out) {  // n_kmers, n_nodes, n_colours, colour_elems, unitig_bases
  Oracle* o = (Oracle*)h; out[0] = o->ix.n_kmers; out[1] = o->ix.nodes.size(); out[2] = o->ix.colours.size();
  u64 ce = 0; for (auto& c : o->ix.colours) ce += c.size(); out[3] = ce;
  u64 ub = 0; for (auto& n : o->ix.nodes) ub += n.seq.size(); out[4] = ub;
}
// Canonical dump of the graph for index parity: per node (sorted by sequence): seq string, colour ids, exts.
// Serialised as text lines "SEQ\tcolour,ids\tlext\trext\n".
u64 orc_index_dump(void* h, char* buf, u64 cap) {
  Oracle* o = (Oracle*)h; std::vector<std::string> lines;
  for (auto& n : o->ix.nodes) {
    std::string l = dna_to_string(n.seq) + "\t";
    const auto& c = o->ix.colours[n.colour];
    for (size_t i = 0; i < c.size(); i++) { if (i) l += ","; l += std::to_string(c[i]); }
    l += "\t" + std::to_string((int)n.lext) + "\t" + std::to_string((int)n.rext) + "\n";
    lines.push_back(l);
  }
  std::sort(lines.begin(), lines.end());
  std::string all; for (auto& l : lines) all += l;
  if (buf && cap >= all.size()) memcpy(buf, all.data(), all.size());
  return all.size();
}

// Run get_calls. scope_off: n_scopes+1 pair offsets (null => one scope over all pairs, FASTQ mode).
// threads: FASTQ mode splits phase 1 over threads only when n_scopes>1 (scopes are independent); a single scope runs
// on one thread like src/process/fastq.rs:15-29 unless `shard_single` != 0, in which case the scope is cut into
// `threads` contiguous shards whose per-shard results are NOT merged (cost model only: used by the all-core baseline).
int orc_run(void* h, const Input* in, u64 n_pairs, const u64* scope_off, u64 n_scopes, int threads, int shard_single) {
  Oracle* o = (Oracle*)h; bool paired = in->r2 != nullptr;
  o->reads.assign(paired ? 2 * n_pairs : n_pairs, ReadRec()); o->pairs.assign(n_pairs, PairRec());
  std::vector<u64> single = {0, n_pairs};
  if (!scope_off) {
    if (shard_single && threads > 1) {
      single.clear(); for (int t = 0; t <= threads; t++) single.push_back(n_pairs * (u64)t / (u64)threads);
      n_scopes = threads;
    } else n_scopes = 1;
    scope_off = single.data();
  }
  o->scopes.assign(n_scopes, ScopeResult()); o->work = Work(); o->error = false;
  MaxinfoTables tables; const MaxinfoTables* tp = nullptr;
  if (in->q1 && !o->cfg.faithful_cost) { maxinfo_tables((size_t)o->cfg.trim_target_length, o->cfg.trim_strictness, tables); tp = &tables; }
  if (threads < 1) threads = 1;
  std::atomic<u64> next(0); std::vector<Work> ws(threads);
  auto worker = [&](int t) {
    for (;;) { u64 s = next.fetch_add(1); if (s >= n_scopes) break; run_scope(*o, *in, scope_off[s], scope_off[s + 1], o->scopes[s], ws[t], tp); }
  };
  std::vector<std::thread> th; for (int t = 1; t < threads; t++) th.emplace_back(worker, t);
  worker(0); for (auto& t : th) t.join();
  for (auto& w : ws) { o->work.probes += w.probes; o->work.nodes += w.nodes; o->work.bases += w.bases; o->work.colour_elems += w.colour_elems; o->work.reads += w.reads; o->work.in_bases += w.in_bases; }
  return o->error ? -1 : 0;
}
void orc_work(void* h, u64* out) { Oracle* o = (Oracle*)h; out[0] = o->work.probes; out[1] = o->work.nodes; out[2] = o->work.bases; out[3] = o->work.colour_elems; out[4] = o->work.reads; out[5] = o->work.in_bases; }

// per read: reason u8, pass u8, score u32, mm u32, trimmed_len u32, ec_len u32 -> 5 x u32 per read; ec ids concatenated separately
u64 orc_read_count(void* h) { return ((Oracle*)h)->reads.size(); }
u64 orc_read_ec_total(void* h) { u64 t = 0; for (auto& r : ((Oracle*)h)->reads) t += r.ec.size(); return t; }
void orc_read_records(void* h, u32* rec, u32* ec) {
  Oracle* o = (Oracle*)h; u64 k = 0;
  for (size_t i = 0; i < o->reads.size(); i++) {
    const ReadRec& r = o->reads[i];
    rec[5 * i] = (u32)r.reason | ((u32)r.pass << 8); rec[5 * i + 1] = r.score; rec[5 * i + 2] = r.mm; rec[5 * i + 3] = r.trimmed_len; rec[5 * i + 4] = (u32)r.ec.size();
    for (u32 e : r.ec) ec[k++] = e;
  }
}
// per pair: triage u8 | fr1<<8 | fr2<<16 | counted<<24, callset index (i32, scope-local)
void orc_pair_records(void* h, u32* rec) {
  Oracle* o = (Oracle*)h;
  for (size_t i = 0; i < o->pairs.size(); i++) { const PairRec& p = o->pairs[i]; rec[2 * i] = (u32)p.triage | ((u32)p.fr1 << 8) | ((u32)p.fr2 << 16) | ((u32)p.counted << 24); rec[2 * i + 1] = (u32)p.callset; }
}
// Results serialised as text: per scope "#scope <i> <n_pairs>\n" then "feat1\tfeat2...\tcount\n" rows (TSV body of src/utils.rs:44-50).
u64 orc_results(void* h, char* buf, u64 cap) {
  Oracle* o = (Oracle*)h; std::string& s = o->blob; s.clear();
  for (size_t i = 0; i < o->scopes.size(); i++) {
    s += "#scope " + std::to_string(i) + " " + std::to_string(o->scopes[i].n_pairs) + "\n";
    for (auto& kv : o->scopes[i].counts) { for (auto& f : kv.first) { s += f; s += "\t"; } s += std::to_string(kv.second); s += "\n"; }
  }
  if (buf && cap >= s.size()) memcpy(buf, s.data(), s.size());
  return s.size();
}

// ---- unit-level entry points used to pin the restatement against the reference's unit tests
double orc_shannon_entropy(const char* s) { return shannon_entropy(std::string(s)); }
u64 orc_maxinfo(const u8* q, u64 n, u64 target, double strictness) { return maxinfo(q, n, target, strictness); }
int orc_natural_lexical_cmp(const char* a, const char* b) { return natural_lexical_cmp(a, b); }
int orc_filter_pair(const u32* a, u64 na, const u32* b, u64 nb) { return filter_pair(std::vector<u32>(a, a + na), std::vector<u32>(b, b + nb)); }
// pseudoalign one read with an explicit min length (src/align.rs:1061-1107 use 12): out = reason, pass, score, mm, ec_len, ec...
void orc_pseudoalign(void* h, const char* seq, u64 min_len, u32* out, double* norm) {
  Oracle* o = (Oracle*)h; ReadRec r; Work w; pseudoalign(o->ix, dna_from_ascii(seq, strlen(seq)), o->cfg, min_len, r, w);
  out[0] = r.reason; out[1] = r.pass; out[2] = r.score; out[3] = r.mm; out[4] = (u32)r.ec.size(); for (size_t i = 0; i < r.ec.size() && i < 59; i++) out[5 + i] = r.ec[i];
  *norm = r.norm;
}
// string-level helpers: inputs/outputs are '\n'-joined lists
static Strs split_lines(const char* s) { Strs v; std::string cur; for (const char* p = s; *p; p++) { if (*p == '\n') { v.push_back(cur); cur.clear(); } else cur += *p; } if (!cur.empty()) v.push_back(cur); return v; }
static u64 join_out(const Strs& v, char* buf, u64 cap) { std::string s; for (auto& x : v) { s += x; s += "\n"; } if (buf && cap > s.size()) { memcpy(buf, s.data(), s.size()); buf[s.size()] = 0; } return s.size(); }
u64 orc_feature_list(void* h, const u32* ec, u64 n, int ignore_rollup, char* buf, u64 cap) { Oracle* o = (Oracle*)h; return join_out(feature_list(std::vector<u32>(ec, ec + n), o->lib, o->cfg, ignore_rollup != 0), buf, cap); }
u64 orc_filter_read_calls(const char* cls, char* buf, u64 cap) { return join_out(filter_read_calls_with_orientation(split_lines(cls)), buf, cap); }
// chemistry filter on two call lists; output "A-lines\n--\nB-lines"
u64 orc_filter_chemistry(const char* a, const char* b, int chem, char* buf, u64 cap) {
  std::vector<Call> ps = parse_calls(split_lines(a)), pm = parse_calls(split_lines(b)); Strs oa, ob;
  if (chem == ChemNone) { for (auto& c : ps) oa.push_back(c.first); for (auto& c : pm) ob.push_back(c.first); }
  else if (chem == Unstranded) { std::vector<Call> fa, fb; filter_unstranded(ps, pm, fa, fb); for (auto& c : fa) oa.push_back(c.first); for (auto& c : fb) ob.push_back(c.first); }
  else filter_prime(ps, pm, chem == FivePrime, oa, ob);
  Strs all = oa; all.push_back("--"); all.insert(all.end(), ob.begin(), ob.end()); return join_out(all, buf, cap);
}
u64 orc_parse_calls(const char* a, char* buf, u64 cap) { Strs o; for (auto& c : parse_calls(split_lines(a))) o.push_back(c.first + (c.second ? "\t1" : "\t0")); return join_out(o, buf, cap); }
u64 orc_intersect(const char* a, const char* b, char* buf, u64 cap) { return join_out(at_intersect(split_lines(a), split_lines(b)), buf, cap); }
int orc_unmap(void* h, const char* feats, u32* out) { Oracle* o = (Oracle*)h; std::vector<u32> r; if (!unmap(split_lines(feats), o->lib, nullptr, r)) return -1; for (size_t i = 0; i < r.size(); i++) out[i] = r[i]; return (int)r.size(); }

}  // extern "C"
