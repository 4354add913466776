// Setup-stage kernels: least-squares initialisation of the foreground coefficients
// (tensorize_fg_coeffs, calibration.py:828-913).  tf.linalg.lstsq's default fast path solves the
// normal equations (A^T A) x = A^T b by Cholesky; here the right-hand sides A^T b come out of the fused
// kernel run in init_mode, the Gram matrices are formed per group from the tiled basis, and one CTA per
// group factors and solves in float64.
#pragma once
#include <cuda_runtime.h>

namespace calb2 {

struct GramJob {
  long long gram_off;  // offset (doubles) of the group's [n][n] matrix in the batch buffer
  int grp;
  int n;               // ncomp
  int slot0;
  int nslots;
  int coef0;
  int pad;
};

struct SlotGeom {
  long long a_off;     // float offset of the owning item's (or basis class's) tile 0
  int item_rows;
  int row_in_item;
  int nbls;
  int swz_ft;          // 0: [tile][item_rows][ft] of the streaming path; 32: shared-basis block, 32-channel tiles with the
                       // 128-byte XOR swizzle (16-byte chunk index ^ (row & 7)) of calfit_shared.cuh
};

// float offset of basis element (row k of the slot, channel f) in the tiled basis
__device__ __forceinline__ long long basis_elem_offset(const SlotGeom& sg, int k, int f, int ft) {
  if (sg.swz_ft) {
    const int row = sg.row_in_item + k, fi = f % sg.swz_ft;
    return sg.a_off + ((long long)(f / sg.swz_ft) * sg.item_rows + row) * sg.swz_ft + ((((fi >> 2) ^ (row & 7)) << 2) | (fi & 3));
  }
  return sg.a_off + ((long long)(f / ft) * sg.item_rows + sg.row_in_item + k) * ft + (f % ft);
}

// Gram matrix of the dense design matrix of calibration.py:897: every baseline of a slot repeats the
// slot's rows, so G[k][k'] = sum_slots nbls_slot * sum_f U[k, slot, f] U[k', slot, f].
// One CTA per group, FC channels staged in shared memory per pass.
template <int FC>
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ A, const GramJob* __restrict__ jobs,
                                                   const SlotGeom* __restrict__ geom, double* __restrict__ gram,
                                                   int nfreqs, int ft) {
  extern __shared__ float tile[];  // [n][FC + 1]
  const GramJob jb = jobs[blockIdx.x];
  const int n = jb.n;
  const int npairs = n * (n + 1) / 2;
  double* Gm = gram + jb.gram_off;
  // each thread owns pairs p = tid, tid + 256, ... and keeps running sums in registers across passes
  constexpr int MAXP = 8;  // pairs per thread per sweep
  for (int pbase = 0; pbase < npairs; pbase += 256 * MAXP) {
    double acc[MAXP];
    int pk[MAXP], pk2[MAXP];
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      acc[m] = 0.0;
      const int pidx = pbase + m * 256 + threadIdx.x;
      int k = 0, k2 = 0;
      if (pidx < npairs) {
        // row-major lower triangle: p = k (k + 1) / 2 + k2, k2 <= k
        k = (int)((sqrt(8.0 * (double)pidx + 1.0) - 1.0) * 0.5);
        while (k * (k + 1) / 2 > pidx) --k;
        while ((k + 1) * (k + 2) / 2 <= pidx) ++k;
        k2 = pidx - k * (k + 1) / 2;
      }
      pk[m] = k;
      pk2[m] = k2;
    }
    for (int s = 0; s < jb.nslots; ++s) {
      const SlotGeom sg = geom[jb.slot0 + s];
      const double wgt = (double)sg.nbls;
      for (int f0 = 0; f0 < nfreqs; f0 += FC) {
        __syncthreads();
        for (int e = threadIdx.x; e < n * FC; e += 256) {
          const int k = e / FC, fi = e % FC, f = f0 + fi;
          float v = 0.f;
          if (f < nfreqs) v = A[basis_elem_offset(sg, k, f, ft)];
          tile[k * (FC + 1) + fi] = v;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MAXP; ++m) {
          if (pbase + m * 256 + threadIdx.x < npairs) {
            const float* ra = tile + pk[m] * (FC + 1);
            const float* rb = tile + pk2[m] * (FC + 1);
            float part = 0.f;
#pragma unroll 8
            for (int fi = 0; fi < FC; ++fi) part = fmaf(ra[fi], rb[fi], part);
            acc[m] += wgt * (double)part;
          }
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      const int pidx = pbase + m * 256 + threadIdx.x;
      if (pidx < npairs) {
        Gm[(long long)pk[m] * n + pk2[m]] = acc[m];
        Gm[(long long)pk2[m] * n + pk[m]] = acc[m];
      }
    }
  }
}

// In-place Cholesky factorisation and two-column solve, one CTA per group.
__global__ void __launch_bounds__(256) chol_solve_kernel(const GramJob* __restrict__ jobs, double* __restrict__ gram,
                                                         float* __restrict__ rhs_r, float* __restrict__ rhs_i) {
  const GramJob jb = jobs[blockIdx.x];
  const int n = jb.n;
  if (n == 0) return;
  double* L = gram + jb.gram_off;
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double d = sqrt(L[(long long)j * n + j]);
    __syncthreads();
    if (threadIdx.x == 0) L[(long long)j * n + j] = d;
    for (int i = j + 1 + threadIdx.x; i < n; i += 256) L[(long long)i * n + j] /= d;
    __syncthreads();
    const int m = n - j - 1;  // trailing size
    for (int e = threadIdx.x; e < m * m; e += 256) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) L[(long long)i * n + k] -= L[(long long)i * n + j] * L[(long long)k * n + j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float* rhs = (threadIdx.x == 0 ? rhs_r : rhs_i) + jb.coef0;
    // reuse the strictly upper triangle's first rows? keep it simple: solve in place in registers/local
    // forward substitution L y = b
    double* ycol = L + (long long)n * n + (long long)threadIdx.x * n;  // scratch appended after the matrix
    for (int i = 0; i < n; ++i) {
      double v = (double)rhs[i];
      for (int k = 0; k < i; ++k) v -= L[(long long)i * n + k] * ycol[k];
      ycol[i] = v / L[(long long)i * n + i];
    }
    // back substitution L^T x = y
    for (int i = n - 1; i >= 0; --i) {
      double v = ycol[i];
      for (int k = i + 1; k < n; ++k) v -= L[(long long)k * n + i] * ycol[k];
      ycol[i] = v / L[(long long)i * n + i];
    }
    for (int i = 0; i < n; ++i) rhs[i] = (float)ycol[i];
  }
}

// multi-GPU: collapse the per-CTA partials into one [4] vector before the all-reduce (fixed order)
__global__ void __launch_bounds__(1024, 1) reduce_partials_kernel(const double* __restrict__ partials, int nitems,
                                                                  double* __restrict__ out) {
  __shared__ double sh[3][32];
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
    a += partials[(size_t)i * 4 + 0];
    b += partials[(size_t)i * 4 + 1];
    c += partials[(size_t)i * 4 + 2];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = a;
    sh[1][warp] = b;
    sh[2][warp] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += sh[0][w];
      b += sh[1][w];
      c += sh[2][w];
    }
    out[0] = a;
    out[1] = b;
    out[2] = c;
    out[3] = 0.0;
  }
}

}  // namespace calb2
