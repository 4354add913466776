// HBM read-bandwidth ceiling probes for the fused kernel's access pattern (not part of the product).
//   ldg   : grid-stride LDG.128 read + reduce
//   bulk  : per-CTA contiguous regions streamed tile by tile with cp.async.bulk into NBUF shared buffers
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(b) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(addr), "r"(parity) : "memory"); } while (!done);
}

__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ a, size_t n4, float* out) {
  float s = 0.f;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = a[i], v1 = a[i + stride], v2 = a[i + 2 * stride], v3 = a[i + 3 * stride];
    s += v0.x + v1.y + v2.z + v3.w;
  }
  for (; i < n4; i += stride) s += a[i].x;
  if (s == 123.456f) out[0] = s;
}

// each CTA streams regions [item * region_bytes, +region_bytes) in tiles of tile_bytes; touch = read smem once
template <int NBUF>
__global__ void __launch_bounds__(256, 2) bulk_kernel(const char* __restrict__ a, int nitems, uint32_t tile_bytes, int ntiles, int touch, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* buf = smem + 128;
  const int tid = threadIdx.x;
  float s = 0.f;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const char* base = a + (size_t)item * tile_bytes * ntiles;
    if (tid == 0) {
      for (int b = 0; b < NBUF; ++b) mbar_init(&mbar[b], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      for (int b = 0; b < NBUF && b < ntiles; ++b) { mbar_expect_tx(&mbar[b], tile_bytes); bulk_g2s(buf + (size_t)b * tile_bytes, base + (size_t)b * tile_bytes, tile_bytes, &mbar[b]); }
    }
    __syncthreads();
    int b = 0; uint32_t par = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&mbar[b], par);
      if (touch) {
        const float4* t = reinterpret_cast<const float4*>(buf + (size_t)b * tile_bytes);
        for (uint32_t e = tid; e < tile_bytes / 16; e += 256) { float4 v = t[e]; s += v.x + v.w; }
      }
      __syncthreads();
      if (tid == 0 && j + NBUF < ntiles) { mbar_expect_tx(&mbar[b], tile_bytes); bulk_g2s(buf + (size_t)b * tile_bytes, base + (size_t)(j + NBUF) * tile_bytes, tile_bytes, &mbar[b]); }
      if (++b == NBUF) { b = 0; par ^= 1; }
    }
    __syncthreads();
  }
  if (s == 123.456f) out[0] = s;
}

template <int NBUF>
float run_bulk(const char* a, size_t bytes, uint32_t tile_bytes, int ntiles, int touch, int persistent, float* out) {
  int nitems = (int)(bytes / ((size_t)tile_bytes * ntiles));
  size_t smem = 128 + (size_t)NBUF * tile_bytes;
  CK(cudaFuncSetAttribute(bulk_kernel<NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = persistent ? 296 : nitems;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    bulk_kernel<NBUF><<<grid, 256, smem>>>(a, nitems, tile_bytes, ntiles, touch, out);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  double gb = (double)nitems * tile_bytes * ntiles / 1e9;
  printf("bulk NBUF=%d tile=%u ntiles=%d touch=%d persistent=%d smem=%zu: %.3f ms  %.1f GB/s\n", NBUF, tile_bytes, ntiles, touch, persistent, smem, best, gb / (best * 1e-3));
  return best;
}

int main() {
  size_t bytes = (size_t)24 << 30;
  char* a; float* out;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(a, 0, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mult : {4, 8, 16, 32}) {
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      ldg_kernel<<<148 * mult, 256>>>((const float4*)a, bytes / 16, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    printf("ldg grid=148x%d: %.3f ms  %.1f GB/s\n", mult, best, bytes / 1e9 / (best * 1e-3));
  }
  // cudaMemcpy D2D for reference (read+write)
  {
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      CK(cudaMemcpyAsync(a, a + bytes / 2, bytes / 2, cudaMemcpyDeviceToDevice));
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    printf("memcpy d2d: %.3f ms  %.1f GB/s (read+write)\n", best, bytes / 1e9 / (best * 1e-3));
  }
  for (int touch : {0, 1}) {
    run_bulk<2>(a, bytes, 45056, 32, touch, 0, out);
    run_bulk<2>(a, bytes, 32768, 32, touch, 0, out);
    run_bulk<3>(a, bytes, 32768, 32, touch, 0, out);
    run_bulk<2>(a, bytes, 16384, 32, touch, 0, out);
    run_bulk<4>(a, bytes, 16384, 32, touch, 0, out);
    run_bulk<6>(a, bytes, 16384, 32, touch, 0, out);
    run_bulk<3>(a, bytes, 32768, 32, touch, 1, out);
  }
  return 0;
}
