// roofs.cu — the roofs the probe / walk traffic is quoted against, measured on the device the job runs on.
//
// "HBM bandwidth for C4 (index >> L2) ... L2 bandwidth for C2/C3/C5 ... measure L2 peak on the box".
// The index traffic of the map stage is random 32-byte sectors (one k-mer bucket) and 64-byte walk records, never a
// stream, so the applicable roof is the random-sector GATHER bandwidth of the memory level the index lives in:
//   nb_measure_gather(table_bytes <= ~L2/2)   -> L2 gather roof   (C2: 47 MB index)
//   nb_measure_gather(table_bytes >> L2)      -> HBM gather roof  (C4: 2-10 GB index), reported beside the stream peak
// and the host link is the roof of the end-to-end number:
//   nb_measure_h2d                            -> pinned host -> device copy bandwidth, one or several GPUs at once
// These are diagnostics behind the C ABI (bench.py prints them next to MEASURED_PEAKS.json); nothing on the data path
// calls them.
#include <cuda_runtime.h>

#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "host.hpp"

using namespace nb;

#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(NB_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

namespace {

__device__ __forceinline__ u64 rmix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// every thread issues U independent record loads per iteration (SECT = 32: one 256-bit load, the k-mer bucket probe;
// SECT = 64: two 256-bit loads of one 64-byte line, the walk record) at pseudo-random record indices
template <int SECT, int U>
__global__ void __launch_bounds__(256) k_gather(const u64* __restrict__ tab, u64 n_rec, u32 iters, u64 seed, u64* sink) {
  u64 x = rmix(((u64)blockIdx.x * blockDim.x + threadIdx.x) ^ seed) | 1ULL;
  u64 acc = 0;
  for (u32 it = 0; it < iters; it++) {
    u64 v[U][SECT / 8];
#pragma unroll
    for (int j = 0; j < U; j++) {
      x = x * 6364136223846793005ULL + 1442695040888963407ULL;
      u64 r = __umul64hi(x, n_rec);
      const u64* p = tab + r * (SECT / 8);
      asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[j][0]), "=l"(v[j][1]), "=l"(v[j][2]), "=l"(v[j][3]) : "l"(p));
      if (SECT == 64) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[j][4]), "=l"(v[j][5]), "=l"(v[j][6]), "=l"(v[j][7]) : "l"(p + 4));
    }
#pragma unroll
    for (int j = 0; j < U; j++)
#pragma unroll
      for (int k = 0; k < SECT / 8; k++) acc ^= v[j][k];
  }
  if (acc == 0x0123456789ABCDEFULL) sink[0] = acc;   // keeps the loads alive
}

__global__ void k_fill(u64* p, u64 n) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) p[i] = rmix(i);
}

}  // namespace

extern "C" {

// Random-record gather bandwidth: bytes = records x record_bytes / time, best of `reps` launches after a warm-up launch.
// record_bytes: 32 or 64.  A table that fits L2 measures the L2 gather roof (the warm-up launch makes it resident).
int nb_measure_gather(int device, uint64_t table_bytes, uint32_t record_bytes, uint32_t reps, double* gbs_out) {
  if (!gbs_out || (record_bytes != 32 && record_bytes != 64) || table_bytes < 4096) return fail(NB_ERR_INVALID, "nb_measure_gather: record_bytes must be 32 or 64, table_bytes >= 4096");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return fail(NB_ERR_CUDA, "no usable CUDA device"); }
  CKR(cudaSetDevice(device));
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  u64 n_rec = table_bytes / record_bytes, n_words = n_rec * (record_bytes / 8);
  u64* tab = nullptr; u64* sink = nullptr;
  CKR(cudaMalloc(&tab, n_words * 8 + 64)); if (cudaMalloc(&sink, 8) != cudaSuccess) { cudaFree(tab); return fail(NB_ERR_CUDA, "cudaMalloc"); }
  k_fill<<<sms * 8, 256>>>(tab, n_words);
  const int U = 4; const unsigned blocks = sms * 8; const u32 iters = 256;   // 2048 threads / SM, 4 records in flight each
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (u32 r = 0; r <= (reps ? reps : 3); r++) {
    cudaEventRecord(e0);
    if (record_bytes == 32) k_gather<32, U><<<blocks, 256>>>(tab, n_rec, iters, 0x9E37 + r, sink);
    else k_gather<64, U><<<blocks, 256>>>(tab, n_rec, iters, 0x9E37 + r, sink);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double gbs = (double)blocks * 256 * iters * U * record_bytes / (ms * 1e-3) / 1e9;
    if (r > 0 && gbs > best) best = gbs;   // launch 0 warms the caches / TLB
  }
  cudaError_t e = cudaDeviceSynchronize();
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(tab); cudaFree(sink);
  if (e != cudaSuccess) return fail(NB_ERR_CUDA, std::string("nb_measure_gather: ") + cudaGetErrorString(e));
  *gbs_out = best;
  return NB_OK;
}

// Pinned host -> device copy bandwidth with the n devices copying at the same time (one host thread, pinned buffer and
// stream per device; `bytes` per copy, `reps` copies each).  out_per_device[n] (GB/s each) may be NULL; *aggregate_out =
// total bytes / wall time of the slowest device.  The ceiling of every end-to-end number that ships host buffers.
int nb_measure_h2d(const int* devices, uint32_t n, uint64_t bytes, uint32_t reps, double* out_per_device, double* aggregate_out) {
  if (!devices || n == 0 || n > 64 || bytes == 0 || !aggregate_out) return fail(NB_ERR_INVALID, "nb_measure_h2d: bad argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(NB_ERR_CUDA, "no usable CUDA device"); }
  for (u32 i = 0; i < n; i++) if (devices[i] < 0 || devices[i] >= ndev) return fail(NB_ERR_INVALID, "nb_measure_h2d: device ordinal out of range");
  if (!reps) reps = 8;
  struct Slot { void* h = nullptr; void* d = nullptr; cudaStream_t s = nullptr; double sec = 0; cudaError_t err = cudaSuccess; };
  std::vector<Slot> sl(n);
  for (u32 i = 0; i < n; i++) {
    Slot& S = sl[i];
    if ((S.err = cudaSetDevice(devices[i])) != cudaSuccess) break;
    if ((S.err = cudaMallocHost(&S.h, bytes)) != cudaSuccess) break;
    if ((S.err = cudaMalloc(&S.d, bytes)) != cudaSuccess) break;
    if ((S.err = cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking)) != cudaSuccess) break;
    memset(S.h, 0x41, bytes);
    S.err = cudaMemcpyAsync(S.d, S.h, bytes, cudaMemcpyHostToDevice, S.s);   // warm-up (page tables, first touch)
    if (S.err == cudaSuccess) S.err = cudaStreamSynchronize(S.s);
  }
  bool ok = true; for (auto& S : sl) ok = ok && S.err == cudaSuccess;
  double wall = 0;
  if (ok) {
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (u32 i = 0; i < n; i++) th.emplace_back([&, i]() {
      Slot& S = sl[i];
      cudaSetDevice(devices[i]);
      auto a = std::chrono::steady_clock::now();
      for (u32 r = 0; r < reps && S.err == cudaSuccess; r++) S.err = cudaMemcpyAsync(S.d, S.h, bytes, cudaMemcpyHostToDevice, S.s);
      if (S.err == cudaSuccess) S.err = cudaStreamSynchronize(S.s);
      S.sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    });
    for (auto& t : th) t.join();
    wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  std::string msg;
  for (u32 i = 0; i < n; i++) {
    Slot& S = sl[i];
    if (S.err != cudaSuccess && msg.empty()) msg = cudaGetErrorString(S.err);
    cudaSetDevice(devices[i]);
    if (S.s) cudaStreamDestroy(S.s);
    if (S.d) cudaFree(S.d);
    if (S.h) cudaFreeHost(S.h);
    if (out_per_device) out_per_device[i] = S.sec > 0 ? (double)bytes * reps / S.sec / 1e9 : 0.0;
  }
  cudaGetLastError();
  if (!msg.empty()) return fail(NB_ERR_CUDA, "nb_measure_h2d: " + msg);
  *aggregate_out = (double)bytes * reps * n / wall / 1e9;
  return NB_OK;
}

}  // extern "C"
