// khash.h — hash of a 60-bit k-mer for the open-addressed k-mer table; shared by the host builder and the device probe.
// Two independent 32-bit hashes from 32-bit multiply / xor-shift only (a 64-bit multiply costs several IMADs per probe
// on the device).
#pragma once
#include <stdint.h>
#ifdef __CUDACC__
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif
NB_HD uint64_t nb_khash(uint64_t km) {
  uint32_t lo = (uint32_t)km, hi = (uint32_t)(km >> 32);
  uint32_t h = (lo * 0x9E3779B1u) ^ (hi * 0x85EBCA77u);
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13;
  uint32_t g = (lo * 0xC2B2AE35u) ^ (hi * 0x27D4EB2Fu); g ^= g >> 15; g *= 0x165667B1u;   // buckets come from the HIGH bits (multiply-high range reduction)
  return (uint64_t)h | ((uint64_t)g << 32);
}
// Two candidate buckets (of two slots each) of the cuckoo k-mer table.  n_buckets is any number in [2, 2^32): the
// 32-bit hashes are range-reduced by multiply-high, so the table can be sized for a chosen load instead of the next
// power of two (at C2 that is a 23 MB table instead of 64 MB: it stays resident in one L2 partition).
NB_HD void nb_cuckoo_buckets(uint64_t km, uint64_t n_buckets, uint32_t& b1, uint32_t& b2) {
  uint64_t h = nb_khash(km);
  b1 = (uint32_t)(((uint64_t)(uint32_t)h * n_buckets) >> 32); b2 = (uint32_t)(((h >> 32) * n_buckets) >> 32);
  if (b2 == b1) b2 = (b1 + 1 == (uint32_t)n_buckets) ? 0u : b1 + 1;
}
// buckets for n distinct k-mers at the build load factor (2 choices x 2 slots: insertion threshold ~0.89)
NB_HD uint64_t nb_cuckoo_size(uint64_t n_kmers, int attempt) {
  uint64_t slots = n_kmers + n_kmers / 3 + 16;          // load 0.75
  for (int i = 0; i < attempt; i++) slots += slots / 4;  // a failed build retries 25 % larger
  return (slots + 1) / 2;
}
