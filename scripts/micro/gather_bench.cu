// gather_bench.cu — what bounds random record gathers on this GPU (round-2 design input for the k-mer table / walk
// record layout).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/gather_bench scripts/micro/gather_bench.cu
// Modes: group G lanes read one random record of G*32 bytes in ONE instruction (G=1: 32 B per lane = a bucket probe;
// G=2: a 64-byte walk record read by a lane pair; G=4: 128 B by a quad), U independent records in flight per thread.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef uint64_t u64; typedef uint32_t u32;
__device__ __forceinline__ u64 rmix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int G, int U, int HINT>
__global__ void __launch_bounds__(256) k(const u64* __restrict__ tab, u64 n_rec, u32 iters, u64 seed, u64* sink) {
  u32 lane = threadIdx.x & 31, sub = lane % G;
  u64 gid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / G;
  u64 x = rmix(gid ^ seed) | 1ULL, acc = 0;
  u64 pol = 0;
  if (HINT == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  if (HINT == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  for (u32 it = 0; it < iters; it++) {
    u64 v[U][4];
#pragma unroll
    for (int j = 0; j < U; j++) {
      x = x * 6364136223846793005ULL + 1442695040888963407ULL;
      u64 r = __umul64hi(x, n_rec);
      const u64* p = tab + r * (4 * G) + 4 * sub;
      if (HINT == 0) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[j][0]), "=l"(v[j][1]), "=l"(v[j][2]), "=l"(v[j][3]) : "l"(p));
      else asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(v[j][0]), "=l"(v[j][1]), "=l"(v[j][2]), "=l"(v[j][3]) : "l"(p), "l"(pol));
    }
#pragma unroll
    for (int j = 0; j < U; j++) acc ^= v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
  }
  if (acc == 0x0123456789ABCDEFULL) sink[0] = acc;
}
__global__ void fill(u64* p, u64 n) { for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = rmix(i); }
template <int G, int U, int HINT> double run(const u64* tab, u64 bytes, int blocks_per_sm, u64* sink) {
  u64 n_rec = bytes / (32 * G); u32 iters = 128; unsigned blocks = 148 * blocks_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); double best = 0;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0); k<G, U, HINT><<<blocks, 256>>>(tab, n_rec, iters, 77 + r, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double g = (double)blocks * 256 * iters * U * 32 / (ms * 1e-3) / 1e9;
    if (r && g > best) best = g;
  }
  return best;
}
int main() {
  u64* sink; cudaMalloc(&sink, 8);
  for (u64 mb : {47ull, 2048ull, 8192ull}) {
    u64 bytes = mb << 20; u64* tab; if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("alloc failed %llu\n", (unsigned long long)mb); continue; }
    fill<<<148 * 8, 256>>>(tab, bytes / 8); cudaDeviceSynchronize();
    printf("table %llu MB: GB/s of 32-byte sectors (G = lanes per record, U = records in flight per thread, bps = blocks of 256 per SM)\n", (unsigned long long)mb);
    printf("  G1 U1 bps8 %.0f | G1 U2 bps8 %.0f | G1 U4 bps8 %.0f | G1 U4 bps4 %.0f | G1 U4 bps2 %.0f | G1 U8 bps4 %.0f\n", run<1, 1, 0>(tab, bytes, 8, sink), run<1, 2, 0>(tab, bytes, 8, sink), run<1, 4, 0>(tab, bytes, 8, sink), run<1, 4, 0>(tab, bytes, 4, sink), run<1, 4, 0>(tab, bytes, 2, sink), run<1, 8, 0>(tab, bytes, 4, sink));
    printf("  G2 U4 bps8 %.0f | G4 U4 bps8 %.0f | G8 U4 bps8 %.0f | G32 U4 bps8 %.0f\n", run<2, 4, 0>(tab, bytes, 8, sink), run<4, 4, 0>(tab, bytes, 8, sink), run<8, 4, 0>(tab, bytes, 8, sink), run<32, 4, 0>(tab, bytes, 8, sink));
    printf("  hints: G1 U4 evict_first %.0f | evict_last %.0f | G2 U4 evict_first %.0f\n", run<1, 4, 1>(tab, bytes, 8, sink), run<1, 4, 2>(tab, bytes, 8, sink), run<2, 4, 1>(tab, bytes, 8, sink));
    cudaFree(tab);
  }
  // mixed: a 64 MB filter read with evict_last while a 2 GB table is gathered with evict_first: does the filter stay in L2?
  return 0;
}
