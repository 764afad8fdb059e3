// Procrustes core of the BASD loss (relational.py:36-50 and its backward) as a Newton-Schulz polar
// iteration on the tensor cores.
//
// Per (extraction point, sample):  C = s_w^T t_w  (D_s x D_t),  loss_b = tr_s + tr_t - 2 ||C||_*.
// The polar factor R of C (C = R H) gives ||C||_* = <R, C>, d||C||_*/ds_w = t_w R^T, d||C||_*/dt_w = s_w R.
// R is never formed: the iterate is kept factored as X_k = W_k t_w with W_k (D_s x N), so that with the weighted,
// centred teacher token Gram K_t = t_w t_w^T (N x N)
//     A_k = X_k X_k^T = W_k K_t W_k^T,      W_{k+1} = (a_k I + b_k A_k + c_k A_k^2) W_k,      W_0 = s_w^T / ||C||_F
// and at convergence   ||C||_* = <K_t W^T, s_w>,   d/ds_w = K_t W^T,   d/dt_w = (s_w W) t_w.
// (a_k, b_k, c_k) are the minimax odd quintics for the shrinking interval [l_k, 1.03] starting at l_0 = 3e-5
// (relative to ||C||_F): every singular value above l_0 ends within 4e-6 of 1 after 10 steps (smaller ones are left
// partially converged - they carry no weight in the nuclear norm); the 3 % head room above 1 keeps rounding from
// pushing the top singular value into the divergent region.
// All products are one-CTA-per-problem tcgen05 GEMMs on split-bf16 operands (polar_gemm.cuh).
// Requires rank(C) = D_s, i.e. D_s <= N - 1 and a teacher token Gram of rank >= D_s.
#include <cstdlib>

#include "cta_linalg.cuh"
#include "polar_gemm.cuh"
#include "spectral.h"

namespace basd {

namespace {

constexpr int kPolarSteps = 10;
constexpr int kPolarChunk = 0;            // problems per launch of the product chain (0 = all); BASD_POLAR_CHUNK overrides
// minimax odd quintics on [l_k, 1.03] (l_0 = 3e-5), each rescaled to a maximum of 1: tools/ns_schedule.py 3e-5 10 1.03
// l_k: 3.0e-5 1.0e-4 4.2e-4 1.7e-3 7.1e-3 2.9e-2 0.118 0.42 0.906 0.99967 -> 0.999996
const float kPolarCoef[kPolarSteps][3] = {
    {4.133071044f, -11.567471663f, 8.093880162f},
    {4.132779328f, -11.565168373f, 8.091968276f},
    {4.131589050f, -11.555734998f, 8.084133962f},
    {4.126671962f, -11.516778686f, 8.051782671f},
    {4.106352567f, -11.356677293f, 7.918925272f},
    {4.022478221f, -10.711657944f, 7.385453943f},
    {3.688471560f, -8.386451534f, 5.490484485f},
    {2.745572830f, -3.677073572f, 1.889197392f},
    {1.941173422f, -1.383242341f, 0.441739912f},
    {1.847826056f, -1.196240433f, 0.348410606f},
};

// ------------------------------------------------------------------------------------------------
// prep_student: s_w = sqrt(a) (s - mu_s)  ->  SW [N][Ds] (split), W_0 = s_w^T [Ds][Np] (split), ksd, tr_s
// one CTA per problem; the whole s_w tile lives in shared memory (fp32, padded rows)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split2(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, float v0, float v1) {   // idx even
    const __nv_bfloat16 h0 = __float2bfloat16(v0), h1 = __float2bfloat16(v1);
    __nv_bfloat162 hv; hv.x = h0; hv.y = h1;
    *reinterpret_cast<__nv_bfloat162*>(hi + idx) = hv;
    *reinterpret_cast<__nv_bfloat162*>(lo + idx) = __floats2bfloat162_rn(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
}

// eight consecutive values -> one 16-byte store per half (idx a multiple of 8)
__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t idx, const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hv);
        const __nv_bfloat162 lv = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        hw[i] = *reinterpret_cast<const uint32_t*>(&hv);
        lw[i] = *reinterpret_cast<const uint32_t*>(&lv);
    }
    *reinterpret_cast<uint4*>(hi + idx) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(lo + idx) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

constexpr int kPrepThreads = 512;
__global__ void __launch_bounds__(kPrepThreads)
polar_prep_student_kernel(PolarArgs g) {
    extern __shared__ float sm[];
    const int N = g.Ns, D = g.Ds, ldS = D + 1;
    float* sw = sm;                                   // [N][D+1]
    float* a_s = sw + static_cast<size_t>(N) * ldS;   // [N]
    float* q_s = a_s + N;
    float* mu = q_s + N;                              // [D]
    float* red = mu + D;                              // 40
    const int prob = blockIdx.x;
    const int i = prob / g.B, b = prob % g.B;
    const __nv_bfloat16* S = g.student[i] + static_cast<size_t>(b) * N * D;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float a = g.a[static_cast<size_t>(prob) * N + n];
        a_s[n] = a;
        q_s[n] = sqrtf(a);
    }
    // raw tokens -> shared memory (bf16x2 loads, D is even)
    for (int t = threadIdx.x; t < N * D / 2; t += blockDim.x) {
        const float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(S)[t]);
        const int n = (2 * t) / D, d = (2 * t) % D;
        sw[n * ldS + d] = v.x;
        sw[n * ldS + d + 1] = v.y;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float m = 0.f;
        for (int n = 0; n < N; ++n) m = fmaf(a_s[n], sw[n * ldS + d], m);
        mu[d] = m;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < N * D; t += blockDim.x) {
        const int n = t / D, d = t % D;
        sw[n * ldS + d] = q_s[n] * (sw[n * ldS + d] - mu[d]);
    }
    __syncthreads();
    float* ksd = g.vec + static_cast<size_t>(prob) * 4 * N;
    float part = 0.f;
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int n = warp; n < N; n += nw) {
            float s = 0.f;
            for (int d = lane; d < D; d += 32) { const float v = sw[n * ldS + d]; s = fmaf(v, v, s); }
            s = warp_sum(s);
            if (lane == 0) { ksd[n] = s; part += s; }
        }
    }
    const float tr_s = cta_sum(part, red);
    if (threadIdx.x == 0) g.scal[prob * 4 + 1] = tr_s;
    // SW = s_w  [N][Ds], tiled [d block][n][64]: pairs of consecutive d
    __nv_bfloat16* swh = g.SW.hi + prob * g.SW.batch_stride;
    __nv_bfloat16* swl = g.SW.lo + prob * g.SW.batch_stride;
    const int Dp = (D + 63) / 64 * 64;
    for (int t = threadIdx.x; t < N * Dp / 2; t += blockDim.x) {
        const int e = 2 * t;                                  // storage index: ((cb * N + n) * 64 + j)
        const int j = e % 64, n = (e / 64) % N, d = (e / (64 * N)) * 64 + j;
        store_split2(swh, swl, e, d < D ? sw[n * ldS + d] : 0.f, d + 1 < D ? sw[n * ldS + d + 1] : 0.f);
    }
    // W_0 = s_w^T [Ds][N], tiled [n block][d][64]: pairs of consecutive n; padding columns stored as zeros
    __nv_bfloat16* wh = g.W.hi + prob * g.W.batch_stride;
    __nv_bfloat16* wl = g.W.lo + prob * g.W.batch_stride;
    const int Np = (N + 63) / 64 * 64;
    for (int t = threadIdx.x; t < D * Np / 2; t += blockDim.x) {
        const int e = 2 * t;                                  // storage index: ((nb * D + d) * 64 + j)
        const int j = e % 64, d = (e / 64) % D, n = (e / (64 * D)) * 64 + j;
        store_split2(wh, wl, e, n < N ? sw[n * ldS + d] : 0.f, n + 1 < N ? sw[(n + 1) * ldS + d] : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// prep_student, D_s a multiple of 8: the raw bf16 tile stays in shared memory as it was read (16-byte loads, 78 KB
// instead of 151 KB of fp32 -> two CTAs per SM overlap each other's load / compute / store phases) and
// s_w = sqrt(a) (s - mu) is recomputed where it is consumed; every global access is 16 bytes wide.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPrepThreads, 2)
polar_prep_student_vec_kernel(PolarArgs g) {
    extern __shared__ __align__(16) uint8_t sm_raw[];
    const int N = g.Ns, D = g.Ds, pitch = D + 8, oct = D / 8;      // pitch in bf16: rows stay 16-byte aligned
    __nv_bfloat16* raw = reinterpret_cast<__nv_bfloat16*>(sm_raw);  // [N][pitch]
    float* a_s = reinterpret_cast<float*>(sm_raw + ((static_cast<size_t>(N) * pitch * 2 + 15) & ~size_t(15)));
    float* q_s = a_s + N;
    float* mu = q_s + N;                                            // [D]
    float* red = mu + D;                                            // 40
    float* partial = red + 40;                                      // [slices][D]
    const int prob = blockIdx.x;
    const int i = prob / g.B, b = prob % g.B;
    const __nv_bfloat16* S = g.student[i] + static_cast<size_t>(b) * N * D;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float a = g.a[static_cast<size_t>(prob) * N + n];
        a_s[n] = a;
        q_s[n] = sqrtf(a);
    }
    for (int t = threadIdx.x; t < N * oct; t += blockDim.x) {
        const int n = t / oct, o = t - n * oct;
        *reinterpret_cast<uint4*>(raw + n * pitch + o * 8) = *reinterpret_cast<const uint4*>(S + static_cast<size_t>(n) * D + o * 8);
    }
    __syncthreads();
    // mu[d] = sum_n a[n] s[n][d]: thread = (octet of d, slice of n)
    const int slices = blockDim.x / oct;
    {
        const int o = threadIdx.x % oct, sl = threadIdx.x / oct;
        if (sl < slices) {
            float m[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int n = sl; n < N; n += slices) {
                const uint4 v = *reinterpret_cast<const uint4*>(raw + n * pitch + o * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
                const float an = a_s[n];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(h[e]);
                    m[2 * e] = fmaf(an, f.x, m[2 * e]);
                    m[2 * e + 1] = fmaf(an, f.y, m[2 * e + 1]);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) partial[sl * D + o * 8 + e] = m[e];
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float m = 0.f;
        for (int sl = 0; sl < slices; ++sl) m += partial[sl * D + d];
        mu[d] = m;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // row norms of s_w
    float* ksd = g.vec + static_cast<size_t>(prob) * 4 * N;
    float part = 0.f;
    for (int n = warp; n < N; n += nw) {
        float s = 0.f;
        const float qn = q_s[n];
        for (int o = lane; o < oct; o += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(raw + n * pitch + o * 8);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                const float x = qn * (f.x - mu[o * 8 + 2 * e]), y = qn * (f.y - mu[o * 8 + 2 * e + 1]);
                s = fmaf(x, x, fmaf(y, y, s));
            }
        }
        s = warp_sum(s);
        if (lane == 0) { ksd[n] = s; part += s; }
    }
    const float tr_s = cta_sum(part, red);
    if (threadIdx.x == 0) g.scal[prob * 4 + 1] = tr_s;
    // SW = s_w [N][Ds], tiled [d block][n][64]: a thread takes 8 consecutive d of one row
    __nv_bfloat16* swh = g.SW.hi + prob * g.SW.batch_stride;
    __nv_bfloat16* swl = g.SW.lo + prob * g.SW.batch_stride;
    const int n_cb = (D + 63) / 64;
    for (int t = threadIdx.x; t < n_cb * N * 8; t += blockDim.x) {
        const int j0 = (t & 7) * 8, n = (t >> 3) % N, cb = (t >> 3) / N;
        const int d0 = cb * 64 + j0;
        float v[8];
        if (d0 < D) {
            const uint4 r = *reinterpret_cast<const uint4*>(raw + n * pitch + d0);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
            const float qn = q_s[n];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[2 * e] = qn * (f.x - mu[d0 + 2 * e]);
                v[2 * e + 1] = qn * (f.y - mu[d0 + 2 * e + 1]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.f;
        }
        store_split8(swh, swl, (static_cast<size_t>(cb) * N + n) * 64 + j0, v);
    }
    // W_0 = s_w^T [Ds][N], tiled [n block][d][64], padding columns zero: a thread takes 8 consecutive n of one d; lanes run
    // along d (conflict-free 2-byte shared reads; the 16-byte stores of a warp land 128 B apart and are merged in L2)
    __nv_bfloat16* wh = g.W.hi + prob * g.W.batch_stride;
    __nv_bfloat16* wl = g.W.lo + prob * g.W.batch_stride;
    const int n_nb = (N + 63) / 64;
    for (int t = threadIdx.x; t < n_nb * 8 * D; t += blockDim.x) {
        const int d = t % D, j0 = ((t / D) & 7) * 8, nb = t / (8 * D);
        const float md = mu[d];
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int n = nb * 64 + j0 + e;
            v[e] = n < N ? q_s[n] * (__bfloat162float(raw[n * pitch + d]) - md) : 0.f;
        }
        store_split8(wh, wl, (static_cast<size_t>(nb) * D + d) * 64 + j0, v);
    }
}

// ------------------------------------------------------------------------------------------------
// prep_teacher: K_t = q (Ktt - m 1^T - 1 m^T + mm) q  (weighted + centred token Gram, split), diag, tr_t
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
polar_prep_teacher_kernel(PolarArgs g) {
    extern __shared__ float sm[];
    const int N = g.Ns;
    float* a_s = sm;                 // [N]
    float* q_s = a_s + N;
    float* m_s = q_s + N;
    float* red = m_s + N;            // 40
    const int prob = blockIdx.x;
    const float* Ktt = g.Ktt + static_cast<size_t>(prob) * N * N;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float a = g.a[static_cast<size_t>(prob) * N + n];
        a_s[n] = a;
        q_s[n] = sqrtf(a);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int n = warp; n < N; n += nw) {
        float s = 0.f;
        for (int m = lane; m < N; m += 32) s = fmaf(Ktt[static_cast<size_t>(n) * N + m], a_s[m], s);
        s = warp_sum(s);
        if (lane == 0) m_s[n] = s;
    }
    __syncthreads();
    float part = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) part += a_s[n] * m_s[n];
    const float mm = cta_sum(part, red);
    float* ktd = g.vec + static_cast<size_t>(prob) * 4 * N + N;
    part = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float v = a_s[n] * (Ktt[static_cast<size_t>(n) * N + n] - 2.f * m_s[n] + mm);
        ktd[n] = v;
        part += v;
    }
    const float tr_t = cta_sum(part, red);
    if (threadIdx.x == 0) g.scal[prob * 4 + 2] = tr_t;
    __nv_bfloat16* kh = g.Kt.hi + prob * g.Kt.batch_stride;
    __nv_bfloat16* kl = g.Kt.lo + prob * g.Kt.batch_stride;
    // storage order [col block][row n][64]: a thread takes 8 consecutive columns of one row (two 16-byte loads of Ktt when
    // N is a multiple of 4, one 16-byte store per half); the second read of Ktt comes from L2
    const int n_cb = (N + 63) / 64;
    const bool vec_ok = (N & 3) == 0;
    for (int t = threadIdx.x; t < n_cb * N * 8; t += blockDim.x) {
        const int j0 = (t & 7) * 8, n = (t >> 3) % N, cb = (t >> 3) / N;
        const int m0 = cb * 64 + j0;
        float v[8];
        const float* row = Ktt + static_cast<size_t>(n) * N;
        if (vec_ok && m0 + 8 <= N) {
            const float4 k0 = *reinterpret_cast<const float4*>(row + m0), k1 = *reinterpret_cast<const float4*>(row + m0 + 4);
            v[0] = k0.x; v[1] = k0.y; v[2] = k0.z; v[3] = k0.w; v[4] = k1.x; v[5] = k1.y; v[6] = k1.z; v[7] = k1.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = m0 + e < N ? row[m0 + e] : 0.f;
        }
        const float qn = q_s[n], mn = m_s[n] - mm;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int m = m0 + e;
            v[e] = m < N ? qn * q_s[m] * (v[e] - mn - m_s[m]) : 0.f;
        }
        store_split8(kh, kl, (static_cast<size_t>(cb) * N + n) * 64 + j0, v);
    }
}

// ------------------------------------------------------------------------------------------------
// finish: nuclear norm, direct student gradient, importance gradient, per-sample loss
//   Gsw = K_t W^T = d nuc / d s_w  [N][Ds]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
polar_finish_kernel(PolarArgs g) {
    extern __shared__ float sm[];
    const int N = g.Ns, D = g.Ds;
    float* dots = sm;                // [N]
    float* ga = dots + N;            // [N]
    float* red = ga + N;             // 40
    const int prob = blockIdx.x;
    const float* G = g.Gsw + static_cast<size_t>(prob) * N * D;
    const __nv_bfloat16* swh = g.SW.hi + prob * g.SW.batch_stride;
    const __nv_bfloat16* swl = g.SW.lo + prob * g.SW.batch_stride;
    const float* a = g.a + static_cast<size_t>(prob) * N;
    float* gdir = g.gdir + static_cast<size_t>(prob) * N * D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if ((D & 7) == 0) {
        // a lane takes 8 consecutive d of one row: 16-byte loads of both halves of s_w, 2 x 16-byte loads of Gsw, 2 x 16-byte
        // stores of gdir; a warp covers 256 columns per pass
        // (the grid is a single wave, so the kernel lasts as long as one warp's chain of rows: four rows are in flight)
        constexpr int RU = 4;
        for (int n0 = warp * RU; n0 < N; n0 += nw * RU) {
            float s[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) s[u] = 0.f;
            for (int d0 = lane * 8; d0 < D; d0 += 256) {
                uint4 h[RU], l[RU];
                float4 g0[RU], g1[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int n = n0 + u < N ? n0 + u : N - 1;
                    const size_t idx = g.SW.at(n, d0);
                    h[u] = *reinterpret_cast<const uint4*>(swh + idx);
                    l[u] = *reinterpret_cast<const uint4*>(swl + idx);
                    g0[u] = *reinterpret_cast<const float4*>(G + static_cast<size_t>(n) * D + d0);
                    g1[u] = *reinterpret_cast<const float4*>(G + static_cast<size_t>(n) * D + d0 + 4);
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    if (n0 + u < N) {
                        const int n = n0 + u;
                        const float qn = sqrtf(a[n]);
                        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h[u]);
                        const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l[u]);
                        float sw[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 x = __bfloat1622float2(hp[e]), y = __bfloat1622float2(lp[e]);
                            sw[2 * e] = x.x + y.x; sw[2 * e + 1] = x.y + y.y;
                        }
                        const float gv[8] = {g0[u].x, g0[u].y, g0[u].z, g0[u].w, g1[u].x, g1[u].y, g1[u].z, g1[u].w};
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) { s[u] = fmaf(sw[e], gv[e], s[u]); o[e] = qn * (2.f * sw[e] - 2.f * gv[e]); }
                        float* od = gdir + static_cast<size_t>(n) * D + d0;
                        *reinterpret_cast<float4*>(od) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(od + 4) = make_float4(o[4], o[5], o[6], o[7]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const float t = warp_sum(s[u]);
                if (lane == 0 && n0 + u < N) dots[n0 + u] = t;
            }
        }
    } else {
        for (int n = warp; n < N; n += nw) {
            const float qn = sqrtf(a[n]);
            float s = 0.f;
            for (int d = lane; d < D; d += 32) {
                const size_t idx = g.SW.at(n, d);
                const float sw = __bfloat162float(swh[idx]) + __bfloat162float(swl[idx]);
                const float gv = G[static_cast<size_t>(n) * D + d];
                s = fmaf(sw, gv, s);
                gdir[static_cast<size_t>(n) * D + d] = qn * (2.f * sw - 2.f * gv);
            }
            s = warp_sum(s);
            if (lane == 0) dots[n] = s;
        }
    }
    __syncthreads();
    const float* ksd = g.vec + static_cast<size_t>(prob) * 4 * N;
    const float* ktd = ksd + N;
    float p_nuc = 0.f, p_gdot = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        p_nuc += dots[n];
        const float v = (ksd[n] + ktd[n] - 2.f * dots[n]) / a[n];
        ga[n] = v;
        p_gdot += v * a[n];
    }
    const float nuc = cta_sum(p_nuc, red);
    const float gdot = cta_sum(p_gdot, red);
    const float inv_ssum = 1.f / g.ssum[prob];
    for (int n = threadIdx.x; n < N; n += blockDim.x) g.gwt[static_cast<size_t>(prob) * N + n] = (ga[n] - gdot) * inv_ssum;
    if (threadIdx.x == 0) {
        const float tr_s = g.scal[prob * 4 + 1], tr_t = g.scal[prob * 4 + 2];
        g.loss_b[prob] = tr_s + tr_t - 2.f * nuc;
        if (g.dbg) {
            g.dbg[prob * 5 + 0] = nuc; g.dbg[prob * 5 + 1] = tr_s; g.dbg[prob * 5 + 2] = tr_t;
            g.dbg[prob * 5 + 3] = static_cast<float>(kPolarSteps); g.dbg[prob * 5 + 4] = g.fro2[prob];
        }
    }
}

}  // namespace

int polar_steps() { return kPolarSteps; }
static bool polar_use_fused(int D, int N) {      // BASD_POLAR_FUSED=0: keep G2 and G3 as two launches (development knob)
    static int fused_cfg = -1;
    if (fused_cfg < 0) {
        const char* e = getenv("BASD_POLAR_FUSED");
        fused_cfg = e ? atoi(e) : 1;
    }
    return fused_cfg != 0 && polar_fused_supported(D, N);
}
int polar_launches_per_step(int Ds, int Ns) { return polar_use_fused(Ds, Ns) ? 3 : 4; }

__device__ long long g_polar_dbg[4][16 * 8];
long long* polar_dbg_ptr(int which) { long long* p = nullptr; cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_polar_dbg); return p + which * 128; }

#define PCK(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t launch_polar_procrustes(const PolarArgs& g, cudaStream_t st, int* launches) {
    const int N = g.Ns, D = g.Ds, nprob = g.n_problems;
    int count = 0;
    {
        TimingScope ts(kSlotPolarPrep, st, 2);
        if (D % 8 == 0) {
            const int slices = kPrepThreads / (D / 8);
            const size_t smem = ((static_cast<size_t>(N) * (D + 8) * 2 + 15) & ~size_t(15)) + (2 * N + D + 40 + static_cast<size_t>(slices) * D) * sizeof(float);
            PCK(cudaFuncSetAttribute(polar_prep_student_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            polar_prep_student_vec_kernel<<<nprob, kPrepThreads, smem, st>>>(g);
        } else {
            const size_t smem = (static_cast<size_t>(N) * (D + 1) + 2 * N + D + 64) * sizeof(float);
            PCK(cudaFuncSetAttribute(polar_prep_student_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            polar_prep_student_kernel<<<nprob, kPrepThreads, smem, st>>>(g);
        }
        PCK(cudaGetLastError());
        polar_prep_teacher_kernel<<<nprob, 256, (3 * N + 64) * sizeof(float), st>>>(g);
        PCK(cudaGetLastError());
        count += 2;
    }
    float* fro2_dense = g.fro2;                // ||C||_F^2 per problem, accumulated by the first A = T W^T
    PCK(cudaMemsetAsync(fro2_dense, 0, sizeof(float) * nprob, st));

    // The chain of products is sequentially dependent per problem but every launch batches many problems.  Problems are
    // walked in chunks small enough for a chunk's matrices (~0.8 MB live per problem) to stay in the 126 MB L2 from one
    // launch to the next; one chunk = one full Newton-Schulz run.  chunk = 0: all problems per launch (HBM streaming).
    static int chunk_cfg = -1;
    if (chunk_cfg < 0) {
        const char* e = getenv("BASD_POLAR_CHUNK");
        chunk_cfg = e ? atoi(e) : kPolarChunk;
    }
    const int chunk = (chunk_cfg <= 0 || chunk_cfg > nprob) ? nprob : chunk_cfg;
    const int n_chunks = (nprob + chunk - 1) / chunk;
    const bool fused = polar_use_fused(D, N);
    auto at = [](const SplitMat& m, int z0) { SplitMat r = m; r.hi += z0 * m.batch_stride; r.lo += z0 * m.batch_stride; return r; };
    TimingScope* gemm_scope = new TimingScope(kSlotPolarGemm, st, n_chunks * ((fused ? 3 : 4) * kPolarSteps + 2));
    struct Del { TimingScope*& p; ~Del() { delete p; } } gemm_del{gemm_scope};
    for (int z0 = 0; z0 < nprob; z0 += chunk) {
        const int nz = nprob - z0 < chunk ? nprob - z0 : chunk;
        // Step 0 runs on the unnormalised W_0 = s_w^T; r = 1 / ||C||_F^2 (trace of W_0 K_t W_0^T, accumulated by the first
        // A product) enters the later epilogues of that step as a per-problem scalar.
        SplitMat Wc = at(g.W, z0), Wn = at(g.W2, z0);
        const SplitMat T = at(g.T, z0), A = at(g.A, z0), Bm = at(g.Bm, z0), Kt = at(g.Kt, z0), SW = at(g.SW, z0);
        float* fro2 = fro2_dense + z0;
        int dir = 0;                           // alternate the problem order launch by launch (L2 reuse of the previous output)
        for (int k = 0; k < kPolarSteps; ++k) {
            const float ca = kPolarCoef[k][0], cb = kPolarCoef[k][1], cc = kPolarCoef[k][2];
            const bool first = k == 0;
            const float* norm = first ? fro2 : nullptr;
            PolarGemmArgs a;
            // G1: T = W K_t
            memset(&a, 0, sizeof a);
            a.epi = PG_EPI_SPLIT; a.out_hi = T.hi; a.out_lo = T.lo; a.out_stride = T.batch_stride; a.scale_c = 1.f;
            if (k == 3 && z0 == 0 && getenv("BASD_POLAR_DBG")) a.dbg_clock = polar_dbg_ptr(0);
            a.reverse = (dir++) & 1;
            PCK(polar_gemm(false, Wc, Kt, nz, a, st));
            if (fused) {
                // G2 + G3 in one kernel: A = T W^T stays in TMEM / shared memory, Bm = a I + b (rA) + c (rA)^2 leaves
                // (step 0: trace(A) = ||C||_F^2 is reduced inside and written to fro2 for G4's epilogue)
                PolarFusedArgs f;
                memset(&f, 0, sizeof f);
                f.ca = ca; f.cb = cb; f.cc = cc; f.first = first ? 1 : 0; f.fro2 = fro2;
                if (k == 3 && z0 == 0 && getenv("BASD_POLAR_DBG")) f.dbg_clock = polar_dbg_ptr(1);
                f.reverse = (dir++) & 1;
                PCK(polar_fused_abm(T, Wc, Bm, nz, f, st));
                count -= 1;
            } else {
                // G2: A = T W^T                (step 0: trace(A) = ||C||_F^2, read by the epilogues of G3 / G4 of that step)
                memset(&a, 0, sizeof a);
                a.epi = PG_EPI_SPLIT; a.out_hi = A.hi; a.out_lo = A.lo; a.out_stride = A.batch_stride; a.scale_c = 1.f;
                if (first) a.trace = fro2;
                if (k == 3 && z0 == 0 && getenv("BASD_POLAR_DBG")) a.dbg_clock = polar_dbg_ptr(1);
                a.reverse = (dir++) & 1;
                PCK(polar_gemm(false, T, Wc, nz, a, st));
                // G3: Bm = a I + b (rA) + c (rA)^2   (A is both operands: the A tile aliases the B tile; the b A term is added from
                //     a TMA-loaded copy of the output-shaped tile of A in the epilogue)
                memset(&a, 0, sizeof a);
                a.epi = PG_EPI_SPLIT; a.out_hi = Bm.hi; a.out_lo = Bm.lo; a.out_stride = Bm.batch_stride;
                a.a_alias_b = 1; a.aux_mode = 1; a.aux_hi = A.hi; a.aux_lo = A.lo;
                a.aux_c = cb; a.aux_p = first ? 1.f : 0.f; a.scale_c = cc; a.scale_p = first ? 2.f : 0.f; a.diag_add = ca; a.norm2 = norm;
                if (k == 3 && z0 == 0 && getenv("BASD_POLAR_DBG")) a.dbg_clock = polar_dbg_ptr(2);
                a.reverse = (dir++) & 1;
                PCK(polar_gemm(false, A, A, nz, a, st));
            }
            // G4: W_next = sqrt(r) Bm W      (W enters as the MN-major B operand; ping-pong buffers)
            memset(&a, 0, sizeof a);
            a.epi = PG_EPI_SPLIT; a.out_hi = Wn.hi; a.out_lo = Wn.lo; a.out_stride = Wn.batch_stride;
            a.scale_c = 1.f; a.scale_p = first ? 0.5f : 0.f; a.norm2 = norm;
            if (k == 3 && z0 == 0 && getenv("BASD_POLAR_DBG")) a.dbg_clock = polar_dbg_ptr(3);
            a.reverse = (dir++) & 1;
            PCK(polar_gemm(true, Bm, Wc, nz, a, st));
            const SplitMat tmp = Wc; Wc = Wn; Wn = tmp;
            count += 4;
        }
        PolarGemmArgs a;
        // Gsw = K_t W^T  [N][Ds]
        memset(&a, 0, sizeof a);
        a.epi = PG_EPI_F32; a.out_f32 = g.Gsw + static_cast<size_t>(z0) * N * D; a.out_f32_stride = static_cast<long long>(N) * D; a.ld_f32 = D;
        a.reverse = (dir++) & 1;
        PCK(polar_gemm(false, Kt, Wc, nz, a, st));
        // Psi = s_w W  [N][N]  -> Theta' = 2 (diag(a) - q Psi q - a a^T), stored as a split pair
        memset(&a, 0, sizeof a);
        a.epi = PG_EPI_THETA; a.out_hi = g.theta + static_cast<size_t>(z0) * N * g.NsPad; a.out_lo = g.theta_lo + static_cast<size_t>(z0) * N * g.NsPad;
        a.out_stride = static_cast<long long>(N) * g.NsPad; a.ld_out = g.NsPad;
        a.vec_a = g.a + static_cast<size_t>(z0) * N;
        a.reverse = (dir++) & 1;
        PCK(polar_gemm(true, SW, Wc, nz, a, st));
        count += 2;
    }
    {
        delete gemm_scope; gemm_scope = nullptr;
        TimingScope tf(kSlotPolarFinish, st, 1);
        polar_finish_kernel<<<nprob, 256, (2 * N + 64) * sizeof(float), st>>>(g);
        PCK(cudaGetLastError());
        count += 1;
    }
    if (launches) *launches = count;
    return cudaSuccess;
}

}  // namespace basd
