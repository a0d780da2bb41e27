// index_build_gpu.cu — K5: coloured compacted stranded de Bruijn index (k = 30) built on the GPU.
//
// Same semantics as the host builder (index_build.cpp) and the same flat output arrays:
//   1  k_enumerate     every 30-mer occurrence -> (k-mer, occurrence index)            one thread per occurrence
//   2  radix sort      by k-mer (cub::DeviceRadixSort, 60 key bits; stable, so sequence ids stay ascending in a group)
//   3  k_group_*       group boundaries -> distinct k-mers, ext masks (OR), colour signature (ordered hash of the ids)
//   4  signature sort  -> colour ids (rank of the signature); representative group per colour -> CSR colour lists
//   5  k_join          succ/pred of the unitig join relation by binary search in the sorted distinct k-mers
//   6  k_walk_chains   one thread per chain start: node / offset of every k-mer; scans give node ids and base offsets
//   7  k_write_unitigs 2-bit unitig store (atomicOr per base word)
//   8  k_table_insert  concurrent insertion into the bucketed k-mer table (atomicCAS on 64-bit keys), then values by lookup
//   9  k_edges         left / right edges by lookup of the neighbouring k-mers
// Universes / bitmap colours (O(colour ids)) are finished on the host (nb_build_universes).  A library with pure
// k-mer cycles (periodic sequence; no chain start) falls back to the host builder, which owns the canonical cycle rule.
// CUB is used for the sort and the scans: the index build is a one-off, not the per-read hot path.
#include <cub/cub.cuh>

#include <cstring>
#include <string>
#include <vector>

#include "host.hpp"
#include "khash.h"

using namespace nb;

namespace {

#define GCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = fail(NB_ERR_CUDA, std::string("index_build_gpu: ") + #x + ": " + cudaGetErrorString(e_)); goto done; } } while (0)

__host__ __device__ __forceinline__ u64 mix64g(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ u64 to_dev_form(u64 be) {   // big-endian 60-bit k-mer -> base i at bits 2i
  u64 x = __brevll(be);
  x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
  return x >> 4;
}
__device__ __forceinline__ u32 upper_seq(const u64* occ_off, u32 n_seq, u64 i) {   // sequence holding occurrence i
  u32 lo = 0, hi = n_seq;
  while (lo < hi) { u32 mid = (lo + hi) >> 1; if (occ_off[mid + 1] <= i) lo = mid + 1; else hi = mid; }
  return lo;
}

__global__ void k_enumerate(const u8* codes, const u64* seq_off, const u64* occ_off, u32 n_seq, u64 n_occ, u64* keys, u32* vals) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n_occ) return;
  u32 s = upper_seq(occ_off, n_seq, i);
  u64 p = seq_off[s] + (i - occ_off[s]);
  u64 km = 0;
  for (int k = 0; k < K; k++) km = (km << 2) | codes[p + k];
  keys[i] = km; vals[i] = (u32)i;
}
__global__ void k_flag_heads(const u64* keys, u64 n, u32* flag) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void k_group_starts(const u32* flag, const u32* gid_excl, u64 n, u64* gstart) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) gstart[gid_excl[i]] = i;
}
// per distinct k-mer: exts, colour signature (same hash as the host builder), number of distinct ids
__global__ void k_group_reduce(const u64* keys, const u32* vals, const u64* gstart, u64 n_groups, u64 n_occ, const u8* codes, const u64* seq_off, const u64* occ_off, u32 n_seq,
                               u64* kmers, u8* Lm, u8* Rm, u64* sig_a, u64* sig_b, u32* n_ids) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  u64 a = gstart[g], b = g + 1 < n_groups ? gstart[g + 1] : n_occ;
  u8 l = 0, r = 0; u64 ha = 0x9E3779B97F4A7C15ULL, hb = 0xC2B2AE3D27D4EB4FULL; u32 last = 0xFFFFFFFFu, cnt = 0;
  for (u64 i = a; i < b; i++) {
    u64 occ = vals[i]; u32 s = upper_seq(occ_off, n_seq, occ);
    u64 pos = occ - occ_off[s], len = seq_off[s + 1] - seq_off[s], p = seq_off[s] + pos;
    if (pos > 0) l |= (u8)(1u << codes[p - 1]);
    if (pos + K < len) r |= (u8)(1u << codes[p + K]);
    if (s != last) { last = s; cnt++; ha = mix64g(ha ^ last) + 0x632BE59BD9B4E019ULL; hb = mix64g(hb + last * 0x9FB21C651E98DF25ULL); }
  }
  kmers[g] = keys[a]; Lm[g] = l; Rm[g] = r; sig_a[g] = ha; sig_b[g] = hb; n_ids[g] = cnt;
}
__global__ void k_sig_heads(const u64* sa, const u64* sb, u64 n, u32* flag) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || sa[i] != sa[i - 1] || sb[i] != sb[i - 1]) ? 1u : 0u;
}
// colour id of every group (rank of its signature) + representative group per colour
__global__ void k_assign_colours(const u32* flag, const u32* cid_excl, const u32* perm, u64 n, u32* col_of_group, u32* rep_group) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 c = cid_excl[i] + flag[i] - 1;      // inclusive scan - 1
  col_of_group[perm[i]] = c;
  if (flag[i]) rep_group[c] = perm[i];
}
__global__ void k_gather64(const u64* src, const u32* idx, u64 n, u64* dst) { u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (i < n) dst[i] = src[idx[i]]; }
__global__ void k_iota(u32* p, u64 n) { u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (i < n) p[i] = (u32)i; }
// colours renumbered by first appearance in k-mer order (what the host builder's interning yields): order[r] = old id
__global__ void k_new_ids(const u32* order, u32 n_col, u32* new_id) { u32 r = blockIdx.x * blockDim.x + threadIdx.x; if (r < n_col) new_id[order[r]] = r; }
__global__ void k_remap(u32* col, u64 n, const u32* new_id) { u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (g < n) col[g] = new_id[col[g]]; }
__global__ void k_colour_sizes(const u32* rep_group, const u32* n_ids, u32 n_col, u32* sizes) {
  u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_col) sizes[c] = n_ids[rep_group[c]];
}
__global__ void k_colour_fill(const u32* rep_group, const u64* gstart, u64 n_groups, u64 n_occ, const u32* vals, const u64* occ_off, u32 n_seq, const u32* col_off, u32 n_col, u32* col_ids) {
  u32 c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  u64 g = rep_group[c], a = gstart[g], b = g + 1 < n_groups ? gstart[g + 1] : n_occ;
  u32 last = 0xFFFFFFFFu, at = col_off[c];
  for (u64 i = a; i < b; i++) { u32 s = upper_seq(occ_off, n_seq, vals[i]); if (s != last) { last = s; col_ids[at++] = s; } }
}
__device__ __forceinline__ u64 find_kmer(const u64* kmers, u64 n, u64 x) {   // index in the sorted distinct k-mers or ~0
  u64 lo = 0, hi = n;
  while (lo < hi) { u64 mid = (lo + hi) >> 1; if (kmers[mid] < x) lo = mid + 1; else hi = mid; }
  return (lo < n && kmers[lo] == x) ? lo : ~0ULL;
}
__global__ void k_join(const u64* kmers, const u8* Lm, const u8* Rm, const u32* col, u64 n, u32* succ, u32* pred) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g >= n) return;
  u8 r = Rm[g];
  if (__popc((unsigned)r) != 1) return;
  u64 y = ((kmers[g] << 2) | (u64)(__ffs((int)r) - 1)) & KMASK;
  u64 j = find_kmer(kmers, n, y);
  if (j != ~0ULL && __popc((unsigned)Lm[j]) == 1 && col[g] == col[j]) { succ[g] = (u32)j; pred[j] = (u32)g; }
}
__global__ void k_start_flags(const u32* pred, u64 n, u32* flag) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g < n) flag[g] = pred[g] == 0xFFFFFFFFu ? 1u : 0u;
}
__global__ void k_walk_chains(const u32* flag, const u32* nid_excl, const u32* succ, u64 n, u32* node_of, u32* off_of, u32* node_first, u32* node_last, u32* node_len) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g >= n || !flag[g]) return;
  u32 id = nid_excl[g]; u64 cur = g; u32 o = 0;
  for (;;) { node_of[cur] = id; off_of[cur] = o; u32 nx = succ[cur]; if (nx == 0xFFFFFFFFu) break; cur = nx; o++; }
  node_first[id] = (u32)g; node_last[id] = (u32)cur; node_len[id] = o + 1;
}
__global__ void k_count_unvisited(const u32* node_of, u64 n, unsigned long long* cnt) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g < n && node_of[g] == 0xFFFFFFFFu) atomicAdd(cnt, 1ULL);
}
__global__ void k_node_bases(const u32* node_len, u32 n_nodes, u64* nb_) { u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n_nodes) nb_[i] = (u64)node_len[i] + K - 1; }
__global__ void k_write_unitigs(const u64* kmers, const u32* succ, const u32* node_first, const u32* node_last, const u32* node_len, const u64* base_start, const u8* Lm, const u8* Rm, const u32* col,
                                u32 n_nodes, unsigned long long* unitig, NodeRec* node) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  u64 pos = base_start[i]; u64 cur = node_first[i]; u64 first = kmers[cur];
  for (int k = K - 1; k >= 0; k--) { u64 bse = (first >> (2 * k)) & 3; atomicOr(unitig + (pos >> 5), (unsigned long long)(bse << (2 * (pos & 31)))); pos++; }
  for (u32 o = 1; o < node_len[i]; o++) { cur = succ[cur]; u64 bse = kmers[cur] & 3; atomicOr(unitig + (pos >> 5), (unsigned long long)(bse << (2 * (pos & 31)))); pos++; }
  NodeRec nr; nr.start_lo = (u32)base_start[i]; nr.len = node_len[i] + K - 1; nr.colour = col[node_first[i]];
  nr.exts_hi = (u32)Lm[node_first[i]] | ((u32)Rm[node_last[i]] << 4) | ((u32)(base_start[i] >> 32) << 8);
  node[i] = nr;
}
// concurrent insertion into the bucketed table (khash.h): first empty slot of the home bucket, else of the next one;
// only keys are placed, values are filled afterwards
__global__ void k_table_insert(const u64* kmers, u64 n, unsigned long long* tkey, u64 nbuckets) {
  u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (g >= n) return;
  u64 dk = to_dev_form(kmers[g]); unsigned long long k = dk | (1ULL << 63);
  u64 b = nb_table_bucket(dk, nbuckets);
  for (;;) {
    for (u64 s = 4 * b; s < 4 * b + 4; s++) if (tkey[s] == 0ULL && atomicCAS(tkey + s, 0ULL, k) == 0ULL) return;
    if (++b == nbuckets) b = 0;
  }
}
__global__ void k_table_values(const unsigned long long* tkey, u64 slots, const u64* kmers, u64 n, const u32* node_of, const u32* off_of, u64* tval) {
  u64 h = blockIdx.x * (u64)blockDim.x + threadIdx.x;
  if (h >= slots) return;
  unsigned long long k = tkey[h];
  if (!(k >> 63)) return;
  // device form -> big-endian to search the sorted array
  u64 d = k & KMASK, be = 0;
  for (int i = 0; i < K; i++) be |= ((d >> (2 * i)) & 3) << (2 * (K - 1 - i));
  u64 g = find_kmer(kmers, n, be);
  tval[h] = (u64)node_of[g] | ((u64)off_of[g] << 32);
}
__global__ void k_edges(const u64* kmers, u64 n, const u32* node_first, const u32* node_last, const u32* node_len, const u8* Lm, const u8* Rm, const u32* node_of, const u32* off_of,
                        u32 n_nodes, u32* redge, u32* ledge, int* bad) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  u64 first = kmers[node_first[i]], last = kmers[node_last[i]]; u8 l = Lm[node_first[i]], r = Rm[node_last[i]];
  for (int b = 0; b < 4; b++) {
    u32 re = 0xFFFFFFFFu, le = 0xFFFFFFFFu;
    if (r >> b & 1) { u64 j = find_kmer(kmers, n, ((last << 2) | (u64)b) & KMASK); if (j == ~0ULL || off_of[j] != 0) *bad = 1; else re = node_of[j]; }
    if (l >> b & 1) { u64 j = find_kmer(kmers, n, (first >> 2) | ((u64)b << 58)); if (j == ~0ULL || off_of[j] != node_len[node_of[j]] - 1) *bad = 1; else le = node_of[j]; }
    redge[4 * (u64)i + b] = re; ledge[4 * (u64)i + b] = le;
  }
}

struct DevMem {   // frees everything on scope exit
  std::vector<void*> p;
  template <class T> cudaError_t alloc(T** out, size_t n) { void* q = nullptr; cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)); if (e == cudaSuccess) { p.push_back(q); *out = (T*)q; } return e; }
  ~DevMem() { for (void* q : p) cudaFree(q); }
};
inline unsigned nblk(u64 n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

int nb_build_index_gpu(const std::vector<std::vector<u8>>& seqs, int device, int n_threads, nb_index** out) {
  int rc = NB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(NB_ERR_CUDA, "no CUDA device available for the GPU index build"); }
  cudaSetDevice(device);
  u32 n_seq = (u32)seqs.size();
  std::vector<u64> seq_off(n_seq + 1, 0), occ_off(n_seq + 1, 0);
  for (u32 s = 0; s < n_seq; s++) { seq_off[s + 1] = seq_off[s] + seqs[s].size(); occ_off[s + 1] = occ_off[s] + (seqs[s].size() >= (size_t)K ? seqs[s].size() - K + 1 : 0); }
  u64 n_bases = seq_off[n_seq], n_occ = occ_off[n_seq];
  if (n_occ == 0 || n_occ >= 0xFFFFFFFFull) return nb_build_index_impl(seqs, n_threads, out);   // empty library, or > 2^32 occurrences: host path
  std::vector<u8> codes(n_bases);
  for (u32 s = 0; s < n_seq; s++) if (!seqs[s].empty()) memcpy(&codes[seq_off[s]], seqs[s].data(), seqs[s].size());
  nb_index* ix = new nb_index(); ix->n_sequences = n_seq;
  DevMem M;
  u8 *d_codes = nullptr, *d_L = nullptr, *d_R = nullptr; u64 *d_seq_off = nullptr, *d_occ_off = nullptr, *d_keys = nullptr, *d_keys2 = nullptr, *d_gstart = nullptr, *d_kmers = nullptr, *d_sa = nullptr, *d_sb = nullptr, *d_sa2 = nullptr, *d_sb2 = nullptr, *d_bases = nullptr, *d_base_start = nullptr, *d_tval = nullptr;
  u32 *d_vals = nullptr, *d_vals2 = nullptr, *d_flag = nullptr, *d_scan = nullptr, *d_nids = nullptr, *d_perm = nullptr, *d_perm2 = nullptr, *d_col = nullptr, *d_rep = nullptr, *d_csz = nullptr, *d_coff = nullptr, *d_cids = nullptr;
  u32 *d_succ = nullptr, *d_pred = nullptr, *d_node_of = nullptr, *d_off_of = nullptr, *d_nfirst = nullptr, *d_nlast = nullptr, *d_nlen = nullptr, *d_redge = nullptr, *d_ledge = nullptr;
  unsigned long long *d_unitig = nullptr, *d_tkey = nullptr, *d_cnt = nullptr; NodeRec* d_node = nullptr; int* d_flagint = nullptr; void* d_tmp = nullptr; size_t tmp_bytes = 0;
  u64 n = 0, slots = 0; u32 n_col = 0, n_nodes = 0; u32 last_scan = 0, last_flag = 0; unsigned long long unvisited = 0; int hflag[2] = {0, 0};
  auto need_tmp = [&](size_t bytes) -> cudaError_t { if (bytes <= tmp_bytes) return cudaSuccess; if (d_tmp) cudaFree(d_tmp); d_tmp = nullptr; tmp_bytes = 0; cudaError_t e = cudaMalloc(&d_tmp, bytes + 256); if (e == cudaSuccess) tmp_bytes = bytes + 256; return e; };
  {
    GCK(M.alloc(&d_codes, n_bases + K)); GCK(M.alloc(&d_seq_off, n_seq + 1)); GCK(M.alloc(&d_occ_off, n_seq + 1));
    GCK(cudaMemcpy(d_codes, codes.data(), n_bases, cudaMemcpyHostToDevice)); GCK(cudaMemcpy(d_seq_off, seq_off.data(), (n_seq + 1) * 8, cudaMemcpyHostToDevice)); GCK(cudaMemcpy(d_occ_off, occ_off.data(), (n_seq + 1) * 8, cudaMemcpyHostToDevice));
    GCK(M.alloc(&d_keys, n_occ)); GCK(M.alloc(&d_keys2, n_occ)); GCK(M.alloc(&d_vals, n_occ)); GCK(M.alloc(&d_vals2, n_occ));
    k_enumerate<<<nblk(n_occ), 256>>>(d_codes, d_seq_off, d_occ_off, n_seq, n_occ, d_keys, d_vals);
    // 2. sort by k-mer
    size_t tb = 0; GCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_keys, d_keys2, d_vals, d_vals2, (int)n_occ, 0, 60)); GCK(need_tmp(tb));
    GCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_keys, d_keys2, d_vals, d_vals2, (int)n_occ, 0, 60));
    // 3. groups
    GCK(M.alloc(&d_flag, n_occ + 1)); GCK(M.alloc(&d_scan, n_occ + 1));
    k_flag_heads<<<nblk(n_occ), 256>>>(d_keys2, n_occ, d_flag);
    tb = 0; GCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_scan, (int)n_occ)); GCK(need_tmp(tb)); GCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_scan, (int)n_occ));
    GCK(cudaMemcpy(&last_scan, d_scan + n_occ - 1, 4, cudaMemcpyDeviceToHost)); GCK(cudaMemcpy(&last_flag, d_flag + n_occ - 1, 4, cudaMemcpyDeviceToHost));
    n = (u64)last_scan + last_flag; ix->n_kmers = n;
    GCK(M.alloc(&d_gstart, n + 1)); k_group_starts<<<nblk(n_occ), 256>>>(d_flag, d_scan, n_occ, d_gstart);
    GCK(M.alloc(&d_kmers, n)); GCK(M.alloc(&d_L, n)); GCK(M.alloc(&d_R, n)); GCK(M.alloc(&d_sa, n)); GCK(M.alloc(&d_sb, n)); GCK(M.alloc(&d_nids, n));
    k_group_reduce<<<nblk(n), 256>>>(d_keys2, d_vals2, d_gstart, n, n_occ, d_codes, d_seq_off, d_occ_off, n_seq, d_kmers, d_L, d_R, d_sa, d_sb, d_nids);
    // 4. colours: sort groups by (sig_b, then sig_a) — two stable passes = lexicographic (sig_a, sig_b)
    GCK(M.alloc(&d_perm, n)); GCK(M.alloc(&d_perm2, n)); GCK(M.alloc(&d_sa2, n)); GCK(M.alloc(&d_sb2, n));
    k_iota<<<nblk(n), 256>>>(d_perm, n);
    tb = 0; GCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_sb, d_sb2, d_perm, d_perm2, (int)n)); GCK(need_tmp(tb));
    GCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_sb, d_sb2, d_perm, d_perm2, (int)n));           // by sig_b; perm2 = order
    // gather sig_a in that order, sort by sig_a (stable), carrying the permutation; then gather both signatures
    {
      // reuse d_keys (n_occ >= n) as gathered sig_a
      u64* ga = d_keys; u64* ga2 = d_sb2;
      k_gather64<<<nblk(n), 256>>>(d_sa, d_perm2, n, ga);
      tb = 0; GCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, ga, ga2, d_perm2, d_perm, (int)n)); GCK(need_tmp(tb));
      GCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, ga, ga2, d_perm2, d_perm, (int)n));              // d_perm = final order
      k_gather64<<<nblk(n), 256>>>(d_sa, d_perm, n, d_sa2); k_gather64<<<nblk(n), 256>>>(d_sb, d_perm, n, d_sb2);
    }
    k_sig_heads<<<nblk(n), 256>>>(d_sa2, d_sb2, n, d_flag);
    tb = 0; GCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_scan, (int)n)); GCK(need_tmp(tb)); GCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_scan, (int)n));
    GCK(cudaMemcpy(&last_scan, d_scan + n - 1, 4, cudaMemcpyDeviceToHost)); GCK(cudaMemcpy(&last_flag, d_flag + n - 1, 4, cudaMemcpyDeviceToHost));
    n_col = last_scan + last_flag;
    GCK(M.alloc(&d_col, n)); GCK(M.alloc(&d_rep, n_col)); GCK(M.alloc(&d_csz, n_col + 1)); GCK(M.alloc(&d_coff, n_col + 1));
    k_assign_colours<<<nblk(n), 256>>>(d_flag, d_scan, d_perm, n, d_col, d_rep);
    { // the head of a signature run is its smallest group index (stable sorts), so sorting colours by d_rep gives first-appearance order
      u32 *d_rep2 = nullptr, *d_ord = nullptr, *d_ord2 = nullptr, *d_newid = nullptr;
      GCK(M.alloc(&d_rep2, n_col)); GCK(M.alloc(&d_ord, n_col)); GCK(M.alloc(&d_ord2, n_col)); GCK(M.alloc(&d_newid, n_col));
      k_iota<<<nblk(n_col), 256>>>(d_ord, n_col);
      tb = 0; GCK(cub::DeviceRadixSort::SortPairs(nullptr, tb, d_rep, d_rep2, d_ord, d_ord2, (int)n_col)); GCK(need_tmp(tb));
      GCK(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_rep, d_rep2, d_ord, d_ord2, (int)n_col));
      k_new_ids<<<nblk(n_col), 256>>>(d_ord2, n_col, d_newid);
      k_remap<<<nblk(n), 256>>>(d_col, n, d_newid);
      d_rep = d_rep2;
    }
    k_colour_sizes<<<nblk(n_col), 256>>>(d_rep, d_nids, n_col, d_csz);
    GCK(cudaMemset(d_csz + n_col, 0, 4));
    tb = 0; GCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_csz, d_coff, (int)n_col + 1)); GCK(need_tmp(tb)); GCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_csz, d_coff, (int)n_col + 1));
    ix->col_off.resize(n_col + 1); GCK(cudaMemcpy(ix->col_off.data(), d_coff, (n_col + 1) * 4, cudaMemcpyDeviceToHost));
    GCK(M.alloc(&d_cids, ix->col_off[n_col]));
    k_colour_fill<<<nblk(n_col), 256>>>(d_rep, d_gstart, n, n_occ, d_vals2, d_occ_off, n_seq, d_coff, n_col, d_cids);
    ix->col_ids.resize(ix->col_off[n_col]); GCK(cudaMemcpy(ix->col_ids.data(), d_cids, (size_t)ix->col_off[n_col] * 4, cudaMemcpyDeviceToHost));
    // 5. join relation
    GCK(M.alloc(&d_succ, n)); GCK(M.alloc(&d_pred, n)); GCK(cudaMemset(d_succ, 0xFF, n * 4)); GCK(cudaMemset(d_pred, 0xFF, n * 4));
    k_join<<<nblk(n), 256>>>(d_kmers, d_L, d_R, d_col, n, d_succ, d_pred);
    // 6. chains
    k_start_flags<<<nblk(n), 256>>>(d_pred, n, d_flag);
    tb = 0; GCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_flag, d_scan, (int)n)); GCK(need_tmp(tb)); GCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_flag, d_scan, (int)n));
    GCK(cudaMemcpy(&last_scan, d_scan + n - 1, 4, cudaMemcpyDeviceToHost)); GCK(cudaMemcpy(&last_flag, d_flag + n - 1, 4, cudaMemcpyDeviceToHost));
    n_nodes = last_scan + last_flag;
    GCK(M.alloc(&d_node_of, n)); GCK(M.alloc(&d_off_of, n)); GCK(cudaMemset(d_node_of, 0xFF, n * 4));
    GCK(M.alloc(&d_nfirst, n_nodes)); GCK(M.alloc(&d_nlast, n_nodes)); GCK(M.alloc(&d_nlen, n_nodes));
    k_walk_chains<<<nblk(n), 256>>>(d_flag, d_scan, d_succ, n, d_node_of, d_off_of, d_nfirst, d_nlast, d_nlen);
    GCK(M.alloc(&d_cnt, 4)); GCK(cudaMemset(d_cnt, 0, 32));
    k_count_unvisited<<<nblk(n), 256>>>(d_node_of, n, d_cnt);
    GCK(cudaMemcpy(&unvisited, d_cnt, 8, cudaMemcpyDeviceToHost));
    if (unvisited) { delete ix; ix = nullptr; if (d_tmp) cudaFree(d_tmp); return nb_build_index_impl(seqs, n_threads, out); }   // pure cycles: the host builder owns the canonical rule
    GCK(M.alloc(&d_bases, n_nodes + 1)); GCK(M.alloc(&d_base_start, n_nodes + 1));
    k_node_bases<<<nblk(n_nodes), 256>>>(d_nlen, n_nodes, d_bases); GCK(cudaMemset(d_bases + n_nodes, 0, 8));
    tb = 0; GCK(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_bases, d_base_start, (int)n_nodes + 1)); GCK(need_tmp(tb)); GCK(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_bases, d_base_start, (int)n_nodes + 1));
    GCK(cudaMemcpy(&ix->unitig_bases, d_base_start + n_nodes, 8, cudaMemcpyDeviceToHost));
    if (ix->unitig_bases >> 40) { rc = fail(NB_ERR_UNSUPPORTED, "unitig store exceeds 2^40 bases"); goto done; }
    // 7. unitig store + node records
    { size_t words = (ix->unitig_bases + 31) / 32 + 2; GCK(M.alloc(&d_unitig, words)); GCK(cudaMemset(d_unitig, 0, words * 8)); GCK(M.alloc(&d_node, n_nodes));
      k_write_unitigs<<<nblk(n_nodes), 256>>>(d_kmers, d_succ, d_nfirst, d_nlast, d_nlen, d_base_start, d_L, d_R, d_col, n_nodes, d_unitig, d_node);
      ix->unitig.resize(words); GCK(cudaMemcpy(ix->unitig.data(), d_unitig, words * 8, cudaMemcpyDeviceToHost));
      ix->node.resize(n_nodes); GCK(cudaMemcpy(ix->node.data(), d_node, (size_t)n_nodes * sizeof(NodeRec), cudaMemcpyDeviceToHost)); }
    // 8. cuckoo table
    GCK(M.alloc(&d_flagint, 2));
    {
      u64 nbk = nb_table_size(n); slots = 4 * nbk;
      if (nbk > 0xFFFFFFFFull) { rc = fail(NB_ERR_UNSUPPORTED, "k-mer table exceeds 2^34 slots"); goto done; }
      GCK(cudaMalloc(&d_tkey, slots * 8)); GCK(cudaMemset(d_tkey, 0, slots * 8));
      k_table_insert<<<nblk(n), 256>>>(d_kmers, n, d_tkey, nbk);
      ix->table_buckets = nbk;
    }
    GCK(M.alloc(&d_tval, slots)); GCK(cudaMemset(d_tval, 0, slots * 8));
    k_table_values<<<nblk(slots), 256>>>(d_tkey, slots, d_kmers, n, d_node_of, d_off_of, d_tval);
    ix->table_key.resize(slots); ix->table_val.resize(slots);
    GCK(cudaMemcpy(ix->table_key.data(), d_tkey, slots * 8, cudaMemcpyDeviceToHost)); GCK(cudaMemcpy(ix->table_val.data(), d_tval, slots * 8, cudaMemcpyDeviceToHost));
    // 9. edges
    GCK(M.alloc(&d_redge, 4 * (size_t)n_nodes)); GCK(M.alloc(&d_ledge, 4 * (size_t)n_nodes)); GCK(cudaMemset(d_flagint, 0, 8));
    k_edges<<<nblk(n_nodes), 256>>>(d_kmers, n, d_nfirst, d_nlast, d_nlen, d_L, d_R, d_node_of, d_off_of, n_nodes, d_redge, d_ledge, d_flagint);
    ix->redge.resize(4 * (size_t)n_nodes); ix->ledge.resize(4 * (size_t)n_nodes);
    GCK(cudaMemcpy(ix->redge.data(), d_redge, ix->redge.size() * 4, cudaMemcpyDeviceToHost)); GCK(cudaMemcpy(ix->ledge.data(), d_ledge, ix->ledge.size() * 4, cudaMemcpyDeviceToHost));
    GCK(cudaMemcpy(hflag, d_flagint, 8, cudaMemcpyDeviceToHost));
    if (hflag[0]) { rc = fail(NB_ERR_INVALID, "internal: de Bruijn edge does not land on a unitig boundary (GPU build)"); goto done; }
    GCK(cudaDeviceSynchronize());
    rc = nb_build_universes(ix, n_seq);
  }
done:
  if (d_tmp) cudaFree(d_tmp);
  if (d_tkey) cudaFree(d_tkey);
  if (rc != NB_OK) { delete ix; return rc; }
  *out = ix;
  return NB_OK;
}

extern "C" int nb_index_build_gpu(const nb_library* lib, int device, int n_threads, nb_index** out) {
  if (!lib || !out) return fail(NB_ERR_INVALID, "null argument");
  const std::vector<std::string>& col = lib->columns[lib->seq_idx];
  std::vector<std::vector<u8>> seqs(col.size());
  for (size_t s = 0; s < col.size(); s++) { seqs[s].resize(col[s].size()); for (size_t i = 0; i < col[s].size(); i++) seqs[s][i] = base_code((u8)col[s][i]); }
  return nb_build_index_gpu(seqs, device, n_threads, out);
}
extern "C" int nb_index_build_gpu_from_sequences(const uint8_t* seq_ascii, const uint64_t* seq_off, uint32_t n_seqs, int device, int n_threads, nb_index** out) {
  if (!seq_off || !out || (!seq_ascii && n_seqs)) return fail(NB_ERR_INVALID, "null argument");
  std::vector<std::vector<u8>> seqs(n_seqs);
  for (u32 s = 0; s < n_seqs; s++) { u64 a = seq_off[s], b = seq_off[s + 1]; seqs[s].resize(b - a); for (u64 i = a; i < b; i++) seqs[s][i - a] = base_code(seq_ascii[i]); }
  return nb_build_index_gpu(seqs, device, n_threads, out);
}
