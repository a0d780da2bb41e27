// synth.cpp — deterministic synthetic workloads of the shapes BASELINE.json names: MHC/KIR-like family
// libraries and 2x150 read pairs / 10x-style single-end records.  Test + bench infrastructure only (not product, not
// oracle).  Counter-based RNG: pair i depends only on (seed, i), so any shard can be produced by any rank or thread.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
typedef uint8_t u8; typedef uint32_t u32; typedef uint64_t u64;
namespace {
struct Rng { u64 s; explicit Rng(u64 a, u64 b) { s = a * 0x9E3779B97F4A7C15ULL ^ (b + 0xD1B54A32D192ED03ULL) * 0xBF58476D1CE4E5B9ULL; next(); }
  u64 next() { u64 z = (s += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }
  double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  u64 below(u64 n) { return next() % n; }
  double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); } };
const char B[5] = "ACGT";
inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; } }
void par(int threads, u64 n, const std::function<void(u64, u64)>& f) { if (threads <= 1 || n < 1024) { f(0, n); return; } std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(f, n * (u64)t / threads, n * (u64)(t + 1) / threads); for (auto& x : th) x.join(); }
}

extern "C" {
// Library: n_fam families x n_all alleles; family base length U[900,1300]; allele = base with SNPs at rate U[0.5%,3%].
// Pass seqs=NULL to get the total size; off has n_fam*n_all+1 entries.
u64 synth_library(u64 seed, u32 n_fam, u32 n_all, char* seqs, u64* off) {
  u64 tot = 0; u64 idx = 0;
  for (u32 f = 0; f < n_fam; f++) {
    Rng r(seed, f); u32 len = 900 + (u32)r.below(401);
    std::vector<char> base(len); for (u32 i = 0; i < len; i++) base[i] = B[r.below(4)];
    for (u32 a = 0; a < n_all; a++) {
      Rng ra(seed ^ 0xA11E1EULL, (u64)f * 64 + a); double rate = 0.005 + 0.025 * ra.uni();
      if (off) off[idx] = tot;
      if (seqs) for (u32 i = 0; i < len; i++) { char c = base[i]; if (a > 0 && ra.uni() < rate) { char d; do d = B[ra.below(4)]; while (d == c); c = d; } seqs[tot + i] = c; }
      tot += len; idx++;
    }
  }
  if (off) off[idx] = tot;
  return tot;
}

struct PairPlan { u32 kind; u32 l1, l2; };  // kind: 0 on-target, 1 off-target, 2 low complexity, 3 short
static u64 source_index(u64 seed, u64 i, double dup_rate) {  // PCR duplicates: pair i repeats an earlier pair exactly
  if (i == 0 || dup_rate <= 0) return i;
  Rng r(seed ^ 0xD0BULL, i); if (r.uni() >= dup_rate) return i; return r.below(i);
}
static PairPlan plan(u64 seed, u64 src, u32 read_len) {
  Rng r(seed, src); double u = r.uni(); PairPlan p; p.l1 = p.l2 = read_len;
  if (u < 0.15) p.kind = 1; else if (u < 0.155) p.kind = 2; else if (u < 0.16) { p.kind = 3; p.l1 = 20 + (u32)r.below(20); p.l2 = 20 + (u32)r.below(20); } else p.kind = 0;
  return p;
}
static void fill_pair(u64 seed, u64 src, u32 read_len, const char* lib, const u64* lib_off, u32 n_seqs, char* o1, char* o2, bool paired, double err) {
  Rng r(seed, src); double u = r.uni(); u32 l1 = read_len, l2 = read_len;
  if (u < 0.15) { for (u32 i = 0; i < l1; i++) o1[i] = B[r.below(4)]; if (paired) for (u32 i = 0; i < l2; i++) o2[i] = B[r.below(4)]; return; }
  if (u < 0.155) { u32 per = 1 + (u32)r.below(2); char m[2] = {B[r.below(4)], B[r.below(4)]}; for (u32 i = 0; i < l1; i++) o1[i] = m[i % per]; if (paired) for (u32 i = 0; i < l2; i++) o2[i] = m[(i + 1) % per]; return; }
  if (u < 0.16) { l1 = 20 + (u32)r.below(20); l2 = 20 + (u32)r.below(20); }
  u32 t = (u32)r.below(n_seqs); u64 t0 = lib_off[t]; u32 tlen = (u32)(lib_off[t + 1] - t0);
  int fl = (int)std::lround(350.0 + 50.0 * r.normal()); u32 need = l1 > l2 ? l1 : l2;
  if (fl < (int)need) fl = (int)need; if ((u32)fl > tlen) fl = (int)tlen;
  u32 st = (u32)r.below((u64)tlen - (u32)fl + 1); bool flip = r.uni() < 0.5;
  // fragment base j (0-based, in fragment orientation)
  auto frag = [&](u32 j) -> char { return flip ? comp(lib[t0 + st + (u32)fl - 1 - j]) : lib[t0 + st + j]; };
  for (u32 i = 0; i < l1; i++) o1[i] = frag(i);
  if (paired) for (u32 i = 0; i < l2; i++) o2[i] = comp(frag((u32)fl - 1 - i));
  auto mutate = [&](char* o, u32 l) { for (u32 i = 0; i < l; i++) if (r.uni() < err) { char d; do d = B[r.below(4)]; while (d == o[i]); o[i] = d; } };
  mutate(o1, l1); if (paired) mutate(o2, l2);
}
// offsets for pairs [first, first+n): r1_off / r2_off have n+1 entries each (relative to this shard)
void synth_pair_offsets(u64 seed, u64 first, u64 n, u32 read_len, double dup_rate, u64* r1_off, u64* r2_off, int threads) {
  std::vector<u32> a(n), b(n);
  par(threads, n, [&](u64 x, u64 y) { for (u64 i = x; i < y; i++) { PairPlan p = plan(seed, source_index(seed, first + i, dup_rate), read_len); a[i] = p.l1; b[i] = p.l2; } });
  r1_off[0] = 0; if (r2_off) r2_off[0] = 0;
  for (u64 i = 0; i < n; i++) { r1_off[i + 1] = r1_off[i] + a[i]; if (r2_off) r2_off[i + 1] = r2_off[i] + b[i]; }
}
void synth_pairs(u64 seed, u64 first, u64 n, u32 read_len, double dup_rate, double err, const char* lib, const u64* lib_off, u32 n_seqs,
                 char* r1, const u64* r1_off, char* r2, const u64* r2_off, int threads) {
  par(threads, n, [&](u64 x, u64 y) { char dummy[2048]; for (u64 i = x; i < y; i++) fill_pair(seed, source_index(seed, first + i, dup_rate), read_len, lib, lib_off, n_seqs, r1 + r1_off[i], r2 ? r2 + r2_off[i] : dummy, r2 != nullptr, err); });
}
// 10x-style single-end records (C3): n UMI groups starting at group `first`; reads/UMI ~ 1 + Geometric(mean 4 total);
// each read L bases from one transcript of the group's gene (forward strand = cDNA sense), 10% of reads get a Q2 tail.
// Outputs: group sizes (n), then per read: bases, raw phred, cell id, scope id (group index).
u64 synth_umi_sizes(u64 seed, u64 first, u64 n, u32* sizes) { u64 tot = 0; for (u64 g = 0; g < n; g++) { Rng r(seed ^ 0x5C0FEULL, first + g); u32 k = 1; while (r.uni() < 0.75 && k < 64) k++; sizes[g] = k; tot += k; } return tot; }
void synth_umi_reads(u64 seed, u64 first, u64 n, const u32* sizes, const u64* read_start, u32 L, u32 n_cells, double err, const char* lib, const u64* lib_off, u32 n_seqs,
                     char* bases, u8* qual, u32* cell, u32* scope, int threads) {
  par(threads, n, [&](u64 x, u64 y) {
    for (u64 g = x; g < y; g++) {
      Rng r(seed ^ 0xBA3ULL, first + g); u32 c = (u32)r.below(n_cells); u32 t = (u32)r.below(n_seqs); bool offt = r.uni() < 0.15;
      u64 t0 = lib_off[t]; u32 tlen = (u32)(lib_off[t + 1] - t0);
      for (u32 k = 0; k < sizes[g]; k++) {
        u64 ri = read_start[g] + k; char* o = bases + ri * L; u8* q = qual + ri * L; cell[ri] = c; scope[ri] = (u32)g;
        bool dup = k > 0 && r.uni() < 0.3;   // PCR duplicate of the previous read of this UMI
        if (dup) { memcpy(o, o - L, L); }
        else if (offt) { for (u32 i = 0; i < L; i++) o[i] = B[r.below(4)]; }
        else { u32 st = (u32)r.below((u64)tlen - L + 1); bool flip = r.uni() < 0.1; for (u32 i = 0; i < L; i++) o[i] = flip ? comp(lib[t0 + st + L - 1 - i]) : lib[t0 + st + i];
               for (u32 i = 0; i < L; i++) if (r.uni() < err) { char d; do d = B[r.below(4)]; while (d == o[i]); o[i] = d; } }
        bool tail = r.uni() < 0.1; u32 tail_at = tail ? 30 + (u32)r.below(L - 30) : L;
        // Phred ~ N(36, 3) approximated by a sum of uniforms (Irwin-Hall, 8 x U[0,255]: sd = 209): no transcendental per base
        for (u32 i = 0; i < L; i++) { int v = 2; if (i < tail_at) { u64 z = r.next(); int sum = 0; for (int k = 0; k < 8; k++) sum += (int)((z >> (8 * k)) & 255); v = 36 + (sum - 1020) * 3 / 209; if (v < 2) v = 2; if (v > 41) v = 41; } q[i] = (u8)v; }
      }
    }
  });
}
}
