// tcgen05 building-block probe for the shared-basis contraction (development aid, not part of the product path).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
//
// Checks, against a float64 host computation, the exact operand forms the fused kernel would use:
//   T1  SS  A K-major SW128 [128 x 32k] . B K-major SW128 [N x 32k]^T          (backward: dC = Q . A^T, Q staged in smem)
//   T2  SS  A K-major SW128 [128 x K]   . B MN-major SW128 [K x 32n]           (forward:  V  = C . A,   C staged in smem)
//   T3  TS  A in TMEM [128 lanes x K cols] . B MN-major                        (forward with C split on the fly into TMEM)
//   T4  TS  A in TMEM . B K-major                                              (backward with dL/dv written to TMEM)
//   T5  3xTF32 split (hi.hi + hi.lo + lo.hi) of T3 at K = 208 with full-precision float inputs: accuracy vs float64
//   T6  the same with the UNTRUNCATED floats as hi operands (does the tensor core drop the low 13 mantissa bits itself?)
// Descriptor fields follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor) and mma_traits_sm100.hpp
// (canonical SW128 layouts: K-major SBO = 8 rows x 128 B, LBO = 1; MN-major with 32 columns and K = 8: no strides used).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e__ = (x);                                                                 \
    if (e__ != cudaSuccess) {                                                              \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);     \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
  return d;
}
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return smem_desc(addr, lbo_bytes, sbo_bytes, 2);
}
// element (row r, column c) of a SWIZZLE_128B_BASE32B tile: 32-byte granules of a 128-byte row XOR-ed with (row & 3)
__device__ __forceinline__ int sw128_b32(int r, int c) { return r * 32 + ((((c >> 3) ^ (r & 3)) << 3) | (c & 7)); }
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// element (row r, 32-bit column c of a 32-column row) of a SW128 tile whose rows are 128 bytes
__device__ __forceinline__ int sw128(int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }

struct Params {
  const float* A;   // [128][K]   row-major (M side: coefficients / dL/dv)
  const float* B;   // T1/T4: [N][32] (rows = N, k contiguous);  T2/T3/T5: [K][32] (rows = k, n contiguous)
  float* D;         // [128][N]
  int K, N, test;
};

// One CTA, 128 threads.  smem: A operand tile(s) + B tile + mbarrier + tmem pointer.
__global__ void __launch_bounds__(128, 1) probe_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* sB = reinterpret_cast<float*>(smem);                // up to 224 rows x 128 B = 28 KB (hi) + 28 KB (lo)
  float* sBlo = sB + 224 * 32;
  float* sA = sBlo + 224 * 32;                               // SS tests: [k-chunk of 32][128 rows x 128 B] = 16 KB per chunk, <= 7 chunks
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int K = p.K, N = p.N;
  const bool b_mn = (p.test == 2 || p.test == 3 || p.test == 5 || p.test == 6);  // B tile is [K rows][32 n] (MN-major operand)
  const bool ts = (p.test >= 3 && p.test <= 6) || p.test == 8;
  const bool k_b32 = p.test == 8;  // K-major operand read out of the MN-major operand's layout (32-byte-base swizzle): one copy for both?
  const bool split = (p.test == 5 || p.test == 6);
  const bool raw_hi = (p.test == 6);  // hi operand = the unmodified float: does the tensor core truncate it to tf32 itself?

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // ---- B tile -> smem (swizzled rows of 32 floats); split: hi and lo copies
  const int brows = b_mn ? K : N;
  for (int e = tid; e < 224 * 32; e += 128) {
    const int r = e >> 5, c = e & 31;
    const float x = r < brows ? p.B[r * 32 + c] : 0.f;
    const float hi = split ? tf32_hi(x) : x;
    const int pos = (b_mn || k_b32) ? sw128_b32(r, c) : sw128(r, c);
    sB[pos] = raw_hi ? x : hi;
    sBlo[pos] = split ? (x - hi) : 0.f;
  }
  // ---- A operand -> smem (SS tests): chunks of 32 k, each [128 rows][32 k] SW128
  if (!ts) {
    for (int e = tid; e < 128 * K; e += 128) {
      const int m = e / K, k = e % K;
      sA[(k >> 5) * 128 * 32 + sw128(m, k & 31)] = p.A[m * K + k];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t COL_D = 0, COL_AHI = 256, COL_ALO = 384;  // D: up to 208 columns; A operand staging (TS): up to 2 x 128... K <= 208 uses chunks
  // ---- TS tests: A -> TMEM.  Row m = thread m.  For K up to 208 the operand is staged chunk by chunk (32 columns each).
  const uint32_t sb_addr = smem_u32(sB), sblo_addr = smem_u32(sBlo), sa_addr = smem_u32(sA);
  const int nchunks = (K + 31) / 32;
  const uint32_t idesc = idesc_tf32(128, b_mn ? 32 : N, 0, b_mn ? 1 : 0);
  uint32_t phase = 0;
  uint32_t accum = 0;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int k0 = ch * 32, kn = min(32, K - k0);  // kn multiple of 8
    if (ts) {
      // stage this chunk's A rows into TMEM columns COL_AHI.. (+ COL_ALO.. for the split)
      for (int c8 = 0; c8 < kn; c8 += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = p.A[tid * K + k0 + c8 + i];
          hi[i] = split ? tf32_hi(x) : x;
          lo[i] = x - hi[i];
          if (raw_hi) hi[i] = x;
        }
        tmem_st8(tmem + lane_base + COL_AHI + c8, hi);
        if (split) tmem_st8(tmem + lane_base + COL_ALO + c8, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncthreads();
    }
    if (tid == 0) {
      tc_fence_after();
      for (int ks = 0; ks < kn; ks += 8) {
        const int k = k0 + ks;
        uint64_t bdesc, bdesc_lo;
        if (b_mn) {  // rows k .. k+7 of the tile = two 4-row atoms of the 32B-base swizzle, 512 B apart (SBO)
          bdesc = smem_desc(sb_addr + (uint32_t)k * 128u, 512, 512, 1);
          bdesc_lo = smem_desc(sblo_addr + (uint32_t)k * 128u, 512, 512, 1);
        } else if (k_b32) {
          bdesc = smem_desc(sb_addr + (uint32_t)k * 4u, 16, 1024, 1);
          bdesc_lo = bdesc;
        } else {     // K-major: N rows, 8 k = 32 B inside the 128 B row
          bdesc = smem_desc_sw128(sb_addr + (uint32_t)k * 4u, 16, 1024);
          bdesc_lo = smem_desc_sw128(sblo_addr + (uint32_t)k * 4u, 16, 1024);
        }
        if (ts) {
          mma_ts(tmem + COL_D, tmem + COL_AHI + ks, bdesc, idesc, accum);
          accum = 1;
          if (split) {
            mma_ts(tmem + COL_D, tmem + COL_AHI + ks, bdesc_lo, idesc, 1);
            mma_ts(tmem + COL_D, tmem + COL_ALO + ks, bdesc, idesc, 1);
          }
        } else {
          const uint64_t adesc = smem_desc_sw128(sa_addr + (uint32_t)ch * 128u * 128u + (uint32_t)ks * 4u, 16, 1024);
          mma_ss(tmem + COL_D, adesc, bdesc, idesc, accum);
          accum = 1;
        }
      }
      mma_commit(&bar);  // arrives when every MMA issued so far has completed (operands may be overwritten, D is final)
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    __syncthreads();
  }
  // ---- epilogue: D -> global
  const int ncols = b_mn ? 32 : N;
  for (int c8 = 0; c8 < ncols; c8 += 8) {
    float v[8];
    tmem_ld8(tmem + lane_base + COL_D + c8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) p.D[tid * ncols + c8 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Cycles per tcgen05.mma (issue to retire, `reps` back to back): form 0 = TS, B MN-major N=32; 1 = SS (A K-major), B MN-major N=32;
// 2 = TS, B K-major N=n; 3 = TS, B MN-major N=64 (two 32-column blocks, LBO); 4-7: accumulator / operand reuse variants
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// (issued behind elect.sync by a converged warp: behind `if (tid == 0)` the compiler wraps every tcgen05.mma in an ELECT retry
// loop and the measurement shows the issuing thread's ~90-170 cycles per MMA, not the tensor pipe)
template <int form>
__global__ void __launch_bounds__(128, 1) time_kernel(int n, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int e = tid; e < 40960; e += 128) reinterpret_cast<float*>(smem)[e] = 0.001f * (e & 15);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
   long long t0 = 0, t1 = 0;
   if (elect_one()) {
    const uint32_t sb = smem_u32(smem);
    const uint64_t b_mn = smem_desc(sb, 8192, 512, 1), b_k = smem_desc_sw128(sb, 16, 1024), a_k = smem_desc_sw128(sb + 65536, 16, 1024);
    const uint32_t id32 = idesc_tf32(128, 32, 0, 1), id64 = idesc_tf32(128, 64, 0, 1), idn = idesc_tf32(128, n, 0, 0);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (form == 0) mma_ts(tmem + 256, tmem + (r & 15) * 8, b_mn + (uint64_t)((r & 7) * 64), id32, 1);
      else if (form == 1) mma_ss(tmem + 256, a_k + (uint64_t)((r & 3) * 2), b_mn + (uint64_t)((r & 7) * 64), id32, 1);
      else if (form == 2) mma_ts(tmem + 256, tmem + (r & 3) * 8, b_k + (uint64_t)((r & 3) * 2), idn, 1);
      else if (form == 3) mma_ts(tmem + 256, tmem + (r & 15) * 8, b_mn + (uint64_t)((r & 7) * 64), id64, 1);
      else if (form == 4) mma_ts(tmem + 256 + 32 * (r & 3), tmem + (r & 15) * 8, b_mn + (uint64_t)((r & 7) * 64), id32, 1);  // 4 accumulators
      else if (form == 5) mma_ts(tmem + 256 + 128 * (r & 1), tmem + (r & 3) * 8, b_k + (uint64_t)((r & 3) * 2), idn, 1);       // 2 accumulators
      else if (form == 6) mma_ts(tmem + 256, tmem + (r & 15) * 8, b_mn + (uint64_t)((r & 7) * 64), id32, r & 7 ? 1 : 0);      // restart every 8
      else mma_ts(tmem + 256, tmem, b_mn, id32, 1);  // same operands every time
    }
    t1 = clock64();
    mma_commit(&bar);
   }
   __syncwarp();
   mbar_wait(&bar, 0);
   const long long t2 = clock64();
   if (t0) {
     out[0] = t1 - t0;
     out[1] = t2 - t0;
   }
  }
  __syncthreads();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  CK(cudaSetDevice(0));
  const int smem_bytes = 2 * 224 * 128 + 7 * 128 * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  struct Case { int test, K, N; const char* name; };
  const Case cases[] = {{1, 32, 208, "T1 SS  A K-major . B K-major   (K=32, N=208)"},
                        {2, 64, 32, "T2 SS  A K-major . B MN-major  (K=64, N=32)"},
                        {3, 64, 32, "T3 TS  A in TMEM . B MN-major  (K=64, N=32)"},
                        {4, 32, 208, "T4 TS  A in TMEM . B K-major   (K=32, N=208)"},
                        {3, 208, 32, "T3 TS  A in TMEM . B MN-major  (K=208, N=32), tf32-exact inputs"},
                        {5, 208, 32, "T5 TS  3xTF32 split, full float inputs (K=208, N=32)"},
                        {6, 208, 32, "T6 TS  3xTF32 split, hi operands = raw floats (hardware truncation) (K=208, N=32)"},
                        {1, 32, 104, "T1 SS  K-major / K-major (K=32, N=104)"},
                        {1, 32, 32, "T1 SS  K-major / K-major (K=32, N=32)"},
                        {2, 8, 32, "T2 SS  A K-major . B MN-major  (K=8, N=32)"},
                        {8, 8, 128, "T8 TS  B K-major read from the 32B-base-swizzled (MN-major) layout (K=8, N=128)"},
                        {8, 32, 208, "T8 TS  B K-major read from the 32B-base-swizzled (MN-major) layout (K=32, N=208)"},
                        {8, 32, 128, "T8 TS  B K-major read from the 32B-base-swizzled (MN-major) layout (K=32, N=128)"}};
  int bad = 0;
  for (const Case& cs : cases) {
    const int K = cs.K, N = cs.N;
    const bool b_mn = (cs.test == 2 || cs.test == 3 || cs.test == 5 || cs.test == 6);
    const int brows = b_mn ? K : N;
    std::vector<float> A(128 * K), B(brows * 32);
    srand(1234 + cs.test + K);
    for (auto& x : A) x = (float)rand() / RAND_MAX - 0.5f;
    for (auto& x : B) x = (float)rand() / RAND_MAX - 0.5f;
    if (cs.test != 5 && cs.test != 6) {  // tf32-exact inputs: the product is then exact up to the accumulation
      for (auto& x : A) x = tf32_trunc(x);
      for (auto& x : B) x = tf32_trunc(x);
    }
    const int ncols = b_mn ? 32 : N;
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4));
    CK(cudaMalloc(&dB, B.size() * 4));
    CK(cudaMalloc(&dD, 128 * ncols * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, 128 * ncols * 4));
    Params p{dA, dB, dD, K, N, cs.test};
    probe_kernel<<<1, 128, smem_bytes>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * ncols);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < ncols; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)(b_mn ? B[k * 32 + n] : B[n * 32 + k]);
        maxerr = std::max(maxerr, std::abs(ref - (double)D[m * ncols + n]));
        maxref = std::max(maxref, std::abs(ref));
      }
    const double rel = maxerr / maxref;
    const bool ok = rel < (cs.test >= 5 ? 3e-6 : 2e-6);
    printf("%-72s max|err| %.3e  max|ref| %.3e  rel %.2e  %s\n", cs.name, maxerr, maxref, rel, ok ? "OK" : "MISMATCH");
    if (!ok) {
      ++bad;
      int nz = 0;
      for (float x : D) nz += (x != 0.f);
      printf("    nonzero outputs: %d of %zu; D[0][0..3] = %g %g %g %g; D[5][0..3] = %g %g %g %g\n", nz, D.size(), D[0], D[1], D[2], D[3],
             D[5 * ncols], D[5 * ncols + 1], D[5 * ncols + 2], D[5 * ncols + 3]);
      double r0 = 0, r1 = 0;
      for (int k = 0; k < K; ++k) {
        r0 += (double)A[k] * (double)(b_mn ? B[k * 32 + 0] : B[k]);
        r1 += (double)A[k] * (double)(b_mn ? B[k * 32 + 1] : B[32 + k]);
      }
      printf("    expected D[0][0..1] = %g %g\n", r0, r1);
    }
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
  }
  {
    long long* dout;
    CK(cudaMalloc(&dout, 16));
    typedef void (*TimeFn)(int, int, long long*);
    const TimeFn fns[8] = {time_kernel<0>, time_kernel<1>, time_kernel<2>, time_kernel<3>, time_kernel<4>, time_kernel<5>, time_kernel<6>, time_kernel<7>};
    for (TimeFn f : fns) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 180224));
    struct TCase { int form, n; const char* name; };
    const TCase tcs[] = {{0, 32, "TS  B MN-major  N=32"}, {1, 32, "SS  B MN-major  N=32"}, {3, 64, "TS  B MN-major  N=64"},
                         {2, 16, "TS  B K-major   N=16"}, {2, 64, "TS  B K-major   N=64"}, {2, 128, "TS  B K-major   N=128"},
                         {4, 32, "TS  MN N=32, 4 accum"}, {5, 128, "TS  K N=128, 2 accum"}, {6, 32, "TS  MN N=32, restart/8"},
                         {7, 32, "TS  MN N=32, same ops"}};
    for (const TCase& tc : tcs)
      for (int reps : {64, 512}) {
        fns[tc.form]<<<1, 128, 180224>>>(tc.n, reps, dout);
        CK(cudaDeviceSynchronize());
        long long h[2];
        CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
        printf("timing %-22s reps %4d: issue %6.1f cyc/MMA, issue->retire %6.1f cyc/MMA\n", tc.name, reps, (double)h[0] / reps,
               (double)h[1] / reps);
      }
    cudaFree(dout);
  }
  printf("%s\n", bad ? "PROBE FAILED" : "PROBE PASSED");
  return bad ? 1 : 0;
}
