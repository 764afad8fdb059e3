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

// C(m,n) = sum_k A[k * lda + m] * B[n * ldb + k]  (A contiguous along m, B contiguous along k: the layouts the products of the
// angles kernel have) with 128-bit loads: per block of four k a thread issues four float4 loads of A (one per k) and four of B
// (one per output column) for 64 FMAs; the scalar accessor version above issues 32 four-byte loads for the same work and was
// bound by the load pipe (133 k cycles for a 192 x 57 x 192 product on one SM; 700 cycles per k).
// When there are fewer tiles than threads, KS = 2 or 4 adjacent lanes split the k range of one tile and combine by shuffle
// (fixed order: bitwise repeatable).  Requires M % 4 == 0, lda % 4 == 0, ldb % 4 == 0 and 16-byte aligned A, B.
// All threads of the CTA must call it (full-mask shuffles).
template <class FC>
__device__ __forceinline__ void cta_gemm_mk(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                            FC store) {
    const int tm = M >> 2, tn = (N + 3) >> 2, tiles = tm * tn;
    const int nthr = static_cast<int>(blockDim.x);
    const int ks = (tiles * 4 <= nthr && K >= 64) ? 4 : (tiles * 2 <= nthr && K >= 32) ? 2 : 1;     // lanes per tile
    const int kc = ((K + ks - 1) / ks + 3) & ~3;                                                   // k range per lane, a multiple of 4
    for (int base = 0; base < tiles * ks; base += nthr) {
        const int idx = base + static_cast<int>(threadIdx.x);
        const bool valid = idx < tiles * ks;
        const int t = valid ? idx / ks : 0, kp = idx % ks;
        const int m0 = (t % tm) << 2, n0 = (t / tm) << 2;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        const float* ap = A + m0;
        const float* bp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bp[j] = B + static_cast<size_t>(min(n0 + j, N - 1)) * ldb;     // (columns past N: computed, never stored)
        const int k_lo = valid ? min(kp * kc, K) : 0, k_hi = valid ? min(k_lo + kc, K) : 0;
        int k = k_lo;
#pragma unroll 2
        for (; k + 4 <= k_hi; k += 4) {
            float4 av[4], bv[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) av[kk] = *reinterpret_cast<const float4*>(ap + static_cast<size_t>(k + kk) * lda);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(bp[j] + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float bk[4] = {bv[j].x, bv[j].y, bv[j].z, bv[j].w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[0][j] = fmaf(av[kk].x, bk[kk], acc[0][j]);
                    acc[1][j] = fmaf(av[kk].y, bk[kk], acc[1][j]);
                    acc[2][j] = fmaf(av[kk].z, bk[kk], acc[2][j]);
                    acc[3][j] = fmaf(av[kk].w, bk[kk], acc[3][j]);
                }
            }
        }
        for (; k < k_hi; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(ap + static_cast<size_t>(k) * lda);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float b = bp[j][k];
                acc[0][j] = fmaf(a4.x, b, acc[0][j]); acc[1][j] = fmaf(a4.y, b, acc[1][j]);
                acc[2][j] = fmaf(a4.z, b, acc[2][j]); acc[3][j] = fmaf(a4.w, b, acc[3][j]);
            }
        }
        if (ks > 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[i][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if (ks == 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
                    acc[i][j] = v;
                }
        }
        if (valid && kp == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + j < N) store(m0 + i, n0 + j, acc[i][j]);
        }
    }
}

// S(m,n) = sum_{k<K} H[k][m] V[k][n] + V[k][m] H[k][n]   (symmetric; H, V row-major [K][ld], ld % 4 == 0, 16-byte aligned)
// for 0 <= m, n < N.  4x4 tiles of the upper triangle only, four 128-bit loads per 32 FMAs, both (m,n) and (n,m) stored.
// (The generic cta_gemm with scalar accessor lambdas spent 4 of every 5 instructions on addresses and loads here.)
template <class FC>
__device__ __forceinline__ void cta_gemm_sym2(int N, int K, const float* __restrict__ H, const float* __restrict__ V, int ld, FC store) {
    const int tn = (N + 3) >> 2;
    for (int t = threadIdx.x; t < tn * tn; t += blockDim.x) {
        const int tm_i = t % tn, tn_i = t / tn;
        if (tm_i > tn_i) continue;                          // lower triangle comes from the mirror tile
        const int m0 = tm_i << 2, n0 = tn_i << 2;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 hm = *reinterpret_cast<const float4*>(H + k * ld + m0), vn = *reinterpret_cast<const float4*>(V + k * ld + n0);
            const float4 vm = *reinterpret_cast<const float4*>(V + k * ld + m0), hn = *reinterpret_cast<const float4*>(H + k * ld + n0);
            const float a0[4] = {hm.x, hm.y, hm.z, hm.w}, b0[4] = {vn.x, vn.y, vn.z, vn.w};
            const float a1[4] = {vm.x, vm.y, vm.z, vm.w}, b1[4] = {hn.x, hn.y, hn.z, hn.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a1[i], b1[j], fmaf(a0[i], b0[j], acc[i][j]));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (m0 + i < N && n0 + j < N) {
                    store(m0 + i, n0 + j, acc[i][j]);
                    if (m0 != n0) store(n0 + j, m0 + i, acc[i][j]);
                }
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
// On exit the lower triangle holds L, the strict upper triangle is zeroed.  *bad (shared int) is set if a pivot at or
// below pivot_floor was met (the pivot is then clamped).
// Left-looking by panels of 8 columns, one matrix row per thread (n <= blockDim.x):
//   1. panel(r, 0..7) = A(r, j..j+7) - sum_{k<j} L(r,k) L(j+c,k)   eight accumulators in registers, one coalesced and two
//      broadcast 128-bit shared loads per 8 FMAs, no barrier inside;
//   2. the 8 x 8 diagonal block is eliminated in 8 mini-steps on the register rows, the pivot row going round through a
//      double-buffered 8-float shared buffer (one barrier per mini-step).
// The right-looking rank-1 version this replaces re-read and re-wrote the whole trailing matrix per column (28 MB of
// shared-memory traffic and 384 barriers for n = 192: 0.2-0.4 ms of the pooled eigen-solver); the subtraction order per
// entry (k ascending) is the same.  ld must be a multiple of 4 and A 16-byte aligned.
__device__ inline void cta_cholesky_lower(float* __restrict__ A, int ld, int n, int* bad, float pivot_floor = 0.f) {
    constexpr int PW = 8;
    __shared__ __align__(16) float s_prow[2][PW];
    int buf = 0;
    for (int j = 0; j < n; j += PW) {
        const int r = j + static_cast<int>(threadIdx.x);          // this thread's row
        const bool row_ok = r < n;
        const int pw = min(PW, n - j);
        float v[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) v[c] = (row_ok && c < pw) ? A[(j + c) * ld + r] : 0.f;
        if (row_ok) {
            const float* lrow = A + r;
            const float* lpan = A + j;                            // L(j + c, k) = A[k * ld + j + c]: 8 consecutive floats
            for (int k = 0; k < j; ++k) {
                const float l = lrow[k * ld];
                const float4 p0 = *reinterpret_cast<const float4*>(lpan + k * ld);
                const float4 p1 = *reinterpret_cast<const float4*>(lpan + k * ld + 4);
                v[0] = fmaf(-l, p0.x, v[0]); v[1] = fmaf(-l, p0.y, v[1]); v[2] = fmaf(-l, p0.z, v[2]); v[3] = fmaf(-l, p0.w, v[3]);
                v[4] = fmaf(-l, p1.x, v[4]); v[5] = fmaf(-l, p1.y, v[5]); v[6] = fmaf(-l, p1.z, v[6]); v[7] = fmaf(-l, p1.w, v[7]);
            }
        }
#pragma unroll
        for (int c = 0; c < PW; ++c) {
            if (c < pw) {
                if (static_cast<int>(threadIdx.x) == c) {         // owner of the pivot row j + c publishes its row
#pragma unroll
                    for (int c2 = 0; c2 < PW; ++c2) s_prow[buf][c2] = v[c2];
                }
                __syncthreads();
                float d = s_prow[buf][c];
                if (!(d > pivot_floor)) { d = fmaxf(pivot_floor, 1e-30f); if (threadIdx.x == 0) *bad = 1; }
                const float inv = rsqrtf(d);
                const float lc = v[c] * inv;                      // L(r, j + c)   (the pivot row itself gets sqrt(d))
                v[c] = lc;
#pragma unroll
                for (int c2 = c + 1; c2 < PW; ++c2) v[c2] = fmaf(-lc, s_prow[buf][c2] * inv, v[c2]);
                buf ^= 1;
            }
        }
        if (row_ok) {
#pragma unroll
            for (int c = 0; c < PW; ++c)
                if (c < pw) A[(j + c) * ld + r] = (r >= j + c) ? v[c] : 0.f;      // upper triangle of the diagonal block zeroed here
        }
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int c = warp; c < n; c += nwarps)
        for (int r = lane; r < c; r += 32) A[c * ld + r] = 0.f;
    __syncthreads();
}

// In-place inverse of a lower-triangular matrix held column-major in shared memory: A[c * ld + r] = L(r, c) on entry,
// = L^-1(r, c) on exit (the strict upper triangle must be zero and stays zero).  Row by row: row i of X = L^-1 is
//   X(i, j) = -( sum_{k = j}^{i-1} L(i, k) X(k, j) ) / L(i, i)   for j < i,      X(i, i) = 1 / L(i, i),
// it needs row i of L (not overwritten yet) and rows < i of X (already in place), so X can take L's storage.  Two adjacent
// lanes per column j split the sum over k (even / odd distance from i - 1); the k loop is uniform over the CTA, so L(i, k) is
// a broadcast read.  One barrier per row (all reads of row i of L before any write to it).  Needs 2 n <= blockDim.x.
// The one-thread-per-column forward substitution into GLOBAL memory this replaces (cta_lower_inverse below) took 0.3 ms per
// 196 x 196 problem - 3.5 of the 4 ms vt_prep_teacher spent on the 1024 problems of BASELINE.json configs[3].
__device__ inline void cta_lower_inverse_inplace(float* __restrict__ A, int ld, int n) {
    const int j = static_cast<int>(threadIdx.x) >> 1, half = static_cast<int>(threadIdx.x) & 1;
    const float* xcol = A + static_cast<size_t>(j < n ? j : 0) * ld;      // X(., j): this thread pair's column
    for (int i = 0; i < n; ++i) {
        const float inv_lii = 1.f / A[i * ld + i];
        float s0 = 0.f, s1 = 0.f;
        if (j < i) {
            int k = i - 1 - half;
            for (; k - 2 >= j; k -= 4) {                                  // two independent chains per thread
                s0 = fmaf(A[k * ld + i], xcol[k], s0);
                s1 = fmaf(A[(k - 2) * ld + i], xcol[k - 2], s1);
            }
            if (k >= j) s0 = fmaf(A[k * ld + i], xcol[k], s0);
        }
        float s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        __syncthreads();                                                  // every thread has read row i of L
        if (half == 0) {
            if (j < i) A[j * ld + i] = -s * inv_lii;
            else if (j == i) A[j * ld + i] = inv_lii;
        }
        // (no second barrier: the next row reads row i + 1 of L, untouched so far, and this pair's own column of X - but the
        //  partner lane's write must be visible: same warp, ordered by the shuffle + barrier of the next iteration)
        __syncwarp();
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
