// CTA-cooperative small dense helpers (fp32, CUDA cores) used by the spectral kernels between Jacobi phases.
// Operands are given as accessor lambdas so triangular masks, scalings and transposes cost nothing extra.
#pragma once
#include <cuda_runtime.h>

namespace basd {

// C(m,n) = sum_k a(m,k) * b(k,n) for 0<=m<M, 0<=n<N; store(m, n, value).  Each thread owns 4x4 tiles;
// consecutive threads take consecutive m-tiles so that accessors contiguous in m coalesce.
template <class FA, class FB, class FC>
__device__ __forceinline__ void cta_gemm(int M, int N, int K, FA a, FB b, FC store) {
    const int tm = (M + 3) >> 2, tn = (N + 3) >> 2;
    for (int t = threadIdx.x; t < tm * tn; t += blockDim.x) {
        const int m0 = (t % tm) << 2, n0 = (t / tm) << 2;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const bool full = (m0 + 4 <= M) && (n0 + 4 <= N);
        if (full) {
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                float av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = a(m0 + i, k);
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = b(k, n0 + j);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        } else {
            for (int k = 0; k < K; ++k) {
                float av[4], bv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = (m0 + i < M) ? a(m0 + i, k) : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = (n0 + j < N) ? b(k, n0 + j) : 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (m0 + i < M && n0 + j < N) store(m0 + i, n0 + j, acc[i][j]);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the whole CTA; result valid in every thread.  scratch: >= 33 floats of shared memory.
__device__ __forceinline__ float cta_sum(float v, float* scratch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        float t = lane < nw ? scratch[lane] : 0.f;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// In-place Cholesky (lower) of the SPD matrix stored column-major in shared memory: A[c*ld + r].
// On exit the lower triangle holds L, the strict upper triangle is zeroed.  Returns false-ish flag through
// *bad (shared int) if a non-positive pivot was met (pivot is then clamped).
__device__ inline void cta_cholesky_lower(float* __restrict__ A, int ld, int n, int* bad, float pivot_floor = 0.f) {
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        float d = A[j * ld + j];
        if (!(d > pivot_floor)) { d = fmaxf(pivot_floor, 1e-30f); if (threadIdx.x == 0) *bad = 1; }
        const float inv = rsqrtf(d);
        __syncthreads();
        for (int r = j + threadIdx.x; r < n; r += blockDim.x) A[j * ld + r] *= inv;   // column j (incl. diag -> sqrt(d))
        __syncthreads();
        // trailing update: A[r][c] -= L[r][j] * L[c][j] for j < c <= r
        const int rem = n - j - 1;
        for (int t = threadIdx.x; t < rem * rem; t += blockDim.x) {
            const int c = j + 1 + t / rem, r = j + 1 + t % rem;
            if (r >= c) A[c * ld + r] = fmaf(-A[j * ld + r], A[j * ld + c], A[c * ld + r]);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        const int c = t / n, r = t % n;
        if (r < c) A[c * ld + r] = 0.f;
    }
    __syncthreads();
}

// Inverse of a lower-triangular matrix L (column-major shared, ld) into Linv (column-major, ld_inv; may be
// global memory).  One thread per column of the inverse (forward substitution), n <= blockDim.x assumed strided.
__device__ inline void cta_lower_inverse(const float* __restrict__ L, int ld, int n, float* __restrict__ Linv, int ld_inv) {
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        float* x = Linv + static_cast<size_t>(c) * ld_inv;
        for (int r = 0; r < c; ++r) x[r] = 0.f;
        x[c] = 1.f / L[c * ld + c];
        for (int r = c + 1; r < n; ++r) {
            float s = 0.f;
            for (int k = c; k < r; ++k) s = fmaf(L[k * ld + r], x[k], s);
            x[r] = -s / L[r * ld + r];
        }
    }
    __syncthreads();
}

}  // namespace basd
