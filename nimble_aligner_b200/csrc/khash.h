// khash.h — hash of a 60-bit k-mer for the open-addressed k-mer table; shared by the host builder and the device probe.
// 32-bit multiply / xor-shift only (a 64-bit multiply costs several IMADs per probe on the device); measured probe
// length on a C2-like k-mer set at load 0.42: 1.364 vs 1.362 for a 64-bit murmur finalizer.
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
  uint32_t g = (hi * 0xC2B2AE35u) ^ lo; g ^= g >> 16;
  return (uint64_t)h | ((uint64_t)g << 32);
}
// Two candidate buckets (of two slots each) of the cuckoo k-mer table; bmask = n_buckets - 1 (power of two, < 2^32).
NB_HD void nb_cuckoo_buckets(uint64_t km, uint64_t bmask, uint32_t& b1, uint32_t& b2) {
  uint64_t h = nb_khash(km);
  b1 = (uint32_t)h & (uint32_t)bmask; b2 = (uint32_t)(h >> 32) & (uint32_t)bmask;
  if (b2 == b1) b2 ^= 1u;
}
