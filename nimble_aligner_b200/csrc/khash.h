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
// The k-mer table is open-addressed over buckets of four 8-byte keys = one 32-byte sector.  A k-mer lives in its home
// bucket or, when that is full, in the next bucket with room (linear probing by bucket); a lookup reads the home
// bucket with one 256-bit load and stops at the first bucket that holds the key or has an empty slot, so a MISS —
// the common answer while an off-target read is searched for a seed — costs one sector (1.07 on average at the
// build load 0.4) instead of the two a 2-choice cuckoo table needs.  n_buckets is any number in [1, 2^32): the
// 32-bit hash is range-reduced by multiply-high.
NB_HD uint32_t nb_table_bucket(uint64_t km, uint64_t n_buckets) { return (uint32_t)(((uint64_t)(uint32_t)nb_khash(km) * n_buckets) >> 32); }
NB_HD uint64_t nb_table_size(uint64_t n_kmers) { return (n_kmers * 5 + 7) / 8 + 4; }   // 4 slots per bucket, load 0.4
// ---- device probe table (built at upload from the flat table above): buckets of two keys and their two values
// {key0, key1, value0, value1} = one 32-byte sector, so a hit needs no second load for (unitig, offset); home bucket by the
// same multiply-high reduction of h, linear probing by bucket.  2 slots at load 0.4: 5 % of buckets are full.
NB_HD uint64_t nb_ptab_buckets(uint64_t n_kmers) { return n_kmers + n_kmers / 4 + 2; }
// Blocked Bloom prefilter in front of it when the table cannot live in L2: one 64-bit word per k-mer (chosen by g, the
// high half of nb_khash), k = 2 or 3 bits inside it from h's low bits.
// The bits sit at fixed halves of the word (first and third in the low 32 bits, second in the high 32), so that building
// and testing them is 32-bit shifts only (a 64-bit variable shift is two instructions on the device, and a probe of an
// off-target read does nothing but this test).
NB_HD uint32_t nb_bloom_lo(uint32_t h, uint32_t k) { uint32_t m = 1u << (h & 31); if (k > 2) m |= 1u << ((h >> 12) & 31); return m; }
NB_HD uint32_t nb_bloom_hi(uint32_t h) { return 1u << ((h >> 6) & 31); }
NB_HD uint64_t nb_bloom_bits(uint32_t h, uint32_t k) { return (uint64_t)nb_bloom_lo(h, k) | ((uint64_t)nb_bloom_hi(h) << 32); }
