// Does inbound bulk-copy (TMA) traffic slow down LDS / STS of the SM's own threads?  (not part of the product)
// Every CTA: thread 0 keeps NBUF bulk copies of `tile_bytes` in flight (optional); warps 1..7 time a loop of
// LDS.128 + FFMA (+ one STS.128 per 8 loads) over a private 32 KB shared-memory region with clock64.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}

constexpr int NBUF = 2;
__global__ void __launch_bounds__(256, 2) probe(const char* __restrict__ a, size_t region_bytes, uint32_t tile_bytes, int stream_on,
                                                int iters, unsigned long long* cyc, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
  float* work = reinterpret_cast<float*>(smem + 128);                       // 32 KB private region
  unsigned char* tbuf = smem + 128 + 32768;                                 // NBUF tile buffers
  __shared__ volatile int stop;
  __shared__ int done_cnt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 8192; e += 256) work[e] = (float)e;
  if (tid == 0) {
    stop = 0;
    done_cnt = 0;
    for (int b = 0; b < NBUF; ++b) mbar_init(&mbar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0 && stream_on) {
      const char* base = a + (size_t)blockIdx.x * region_bytes;
      size_t off = 0;
      uint32_t par[NBUF] = {0, 0};
      for (int b = 0; b < NBUF; ++b) { mbar_expect_tx(&mbar[b], tile_bytes); bulk_g2s(tbuf + (size_t)b * tile_bytes, base + off, tile_bytes, &mbar[b]); off += tile_bytes; }
      int b = 0;
      while (!stop) {
        if (mbar_test(&mbar[b], par[b])) {
          par[b] ^= 1;
          if (off + tile_bytes > region_bytes) off = 0;
          mbar_expect_tx(&mbar[b], tile_bytes);
          bulk_g2s(tbuf + (size_t)b * tile_bytes, base + off, tile_bytes, &mbar[b]);
          off += tile_bytes;
          b ^= 1;
        }
      }
      // drain
      for (int q = 0; q < NBUF; ++q) while (!mbar_test(&mbar[q], par[q])) {}
    }
  } else {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* w4 = reinterpret_cast<const float4*>(work);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = w4[((warp - 1) * 8 + k) * 32 + lane];   // conflict-free 512 B per warp
        acc.x = fmaf(v.x, 1.0001f, acc.x); acc.y = fmaf(v.y, 1.0001f, acc.y);
        acc.z = fmaf(v.z, 1.0001f, acc.z); acc.w = fmaf(v.w, 1.0001f, acc.w);
      }
      if (lane < 8) reinterpret_cast<float4*>(work)[1792 + (warp - 1) * 8 + lane] = acc;
    }
    const long long t1 = clock64();
    if (lane == 0) atomicAdd(cyc, (unsigned long long)(t1 - t0));
    if (acc.x == 123.f) sink[0] = acc.y;
  }
  // warps 1..7 done -> stop the streamer
  if (warp != 0 && lane == 0) { __threadfence_block(); atomicAdd((int*)&done_cnt, 1); }
  if (warp != 0 && lane == 0) { while (atomicAdd((int*)&done_cnt, 0) < 7) {} stop = 1; }
}

int main() {
  const size_t region = (size_t)32 << 20;  // 32 MB per CTA
  const int grid = 296;
  char* a; unsigned long long* cyc; float* sink;
  CK(cudaMalloc(&a, region * grid)); CK(cudaMemset(a, 0, region * grid));
  CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 4));
  const int iters = 20000;
  for (uint32_t tile : {16384u, 32768u}) {
    const size_t smem = 128 + 32768 + (size_t)NBUF * tile;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int on : {0, 1, 0, 1}) {
      CK(cudaMemset(cyc, 0, 8));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<<<grid, 256, smem>>>(a, region, tile, on, iters, cyc, sink);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
      const double per_iter = (double)h / (grid * 7.0) / iters;   // cycles per 8 LDS.128 (+1 STS) per warp
      printf("tile %u stream %d: %.1f cycles per 8x LDS.128+STS per warp (7 warps x 2 CTAs/SM busy), kernel %.2f ms\n", tile, on, per_iter, ms);
    }
  }
  return 0;
}
