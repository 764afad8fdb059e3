// One-sided (Hestenes) Jacobi on a column-major fp32 matrix held in shared memory.
//
// Orthogonalises the n columns (length m, leading dimension ld) of A in place: A <- A V with V orthogonal,
// until every pair of columns satisfies |a_p . a_q| <= tol * |a_p| |a_q|.  On exit the column norms are the
// singular values of the input and the normalised columns its left singular vectors; for a symmetric PSD
// input (A = G) the norms are the eigenvalues and the normalised columns the eigenvectors (G v = lambda v),
// so no separate rotation accumulator is needed.  Replaces the LAPACK calls at
// /root/reference/src/losses/layer_selector.py:16,36,92,99 and relational.py:48.
//
// A pair of columns is owned by an 8-lane group: each lane keeps its slice of both columns in registers (128-bit
// shared accesses; a quarter warp touches 128 contiguous bytes, so they are bank-conflict free), the three inner
// products are reduced with 3 xor-shuffles, the rotation is applied from registers in packed fp32.  Two orderings:
// odd-even transposition with register-resident columns over a thread-block cluster (the big pooled problems) and a
// round-robin tournament on one CTA (small problems).  ld must be a multiple of 4.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace basd {

constexpr int JAC_GROUP = 8;
constexpr int JAC_CHUNK_ROWS = JAC_GROUP * 4;
constexpr int JAC_MAX_CHUNKS = 8;      // 8 chunks x 32 rows -> m <= 256 (register-resident variants)

__host__ __device__ inline int jacobi_ld(int m) { return (m + 3) & ~3; }     // 128-bit rows; no bank constraint (see above)

// pair k (0 <= k < n/2) at step s (0 <= s < n-1); n even.
__device__ __forceinline__ void jacobi_pair(int n, int s, int k, int& p, int& q) {
    if (k == 0) { p = n - 1; q = s; return; }
    const int m1 = n - 1;
    p = s + k; if (p >= m1) p -= m1;
    q = s - k; if (q < 0) q += m1;
}

// ------------------------------------------------------------------------------------------------------------------
// Odd-even variant with register-resident columns, spread over a thread-block CLUSTER (the big pooled problems).
// Positions 0..n-1 hold the columns; group g owns positions 2g (registers P) and 2g+1 (registers Q).  A sweep is n
// steps of the odd-even transposition network, every step followed by a swap of the two columns it paired, so that
// after n steps every pair of columns has met exactly once:
//   even step : rotate (P, Q) in registers - no shared memory, no barrier;
//   odd step  : positions (2g+1, 2g+2): group g keeps its odd column, borrows the even column of group g+1 through
//               that group's mailbox (column 2g+2 of A), rotates, keeps the borrowed one and mails the other back.
// The n/2 groups are dealt to the CTAs of the cluster in contiguous runs, so one SM issues the instructions and carries
// the shared-memory traffic of n/(2C) pairs.  Inside a CTA the two barriers per pair-step are a named barrier over the
// warps that hold groups.  Across a CTA boundary (last group g of CTA k, first group g+1 of CTA k+1) the exchange is
// pushed through distributed shared memory with st.async, which credits an mbarrier of the RECEIVING CTA with the bytes
// it delivered:
//   group g+1 pushes its even column into CTA k's inbox   -> bar_in  of CTA k   (instead of mailing it locally)
//   group g   rotates against the inbox and pushes the column that moves on into group g+1's mailbox -> bar_back of CTA k+1
// so no cluster-wide barrier sits in the pair-step (barrier.cluster with release/acquire compiles to MEMBAR.ALL.GPU +
// CCTL.IVALL; with two of them per pair-step the cluster version was slower than one CTA).
// The matrix lives in the shared memory of cluster rank 0 (same offset `A` in every CTA); the other CTAs use their own
// copy of that region for mailboxes only.  One cluster barrier per sweep ORs the convergence flags.
//
// The pair-step is a dependent chain (measured ~1.8k cycles with six warps on an SM, whatever the column length), so
// the chain is what the code below shortens:
//  * every lane of a warp runs the same instruction stream (inactive groups rotate zero columns; memory operations are
//    predicated), so the 8-lane reductions are full-mask shuffles - with per-group masks the compiler routes every
//    shuffle through a divergent WARPSYNC.COLLECTIVE path that serialises the four groups of the warp;
//  * the rotation comes from d = beta - alpha, h = 2 gamma through two dependent rsqrt (cos 2theta = |d| / hypot(d, h)),
//    not from the five-MUFU tan(theta) chain, is computed speculatively and discarded if the pair is already orthogonal;
//  * packed fp32 (FFMA2 / FMUL2) for the dot products and the rotation.
// Column ORDER on exit is a permutation of the input order (irrelevant to the callers, which sort by eigenvalue).
// Measured on the 28 pooled 192 x 192 problems of cfg2 (ms, whole kernel): round-robin 4.7, odd-even on one CTA 4.1,
// + packed fp32 3.8, 4-CTA cluster 2.8, + uniform pair-step / 256-thread CTAs 2.1, + panel Cholesky 1.8, + end-game exit
// 1.5; exchanging in-warp neighbours by shuffle instead of mailboxes was slower (4.7); clusters of 3 and 4 CTAs measure
// the same, 5 do not all fit the GPU at once (2.5).
// All threads of all CTAs of the cluster must call it; returns the sweep count (identical in every CTA).
// inbox: ld floats, bars: 2 mbarriers, flags: 16 ints - shared memory at the same offsets in every CTA.
// ------------------------------------------------------------------------------------------------------------------
using jac_f2 = unsigned long long;          // two fp32 lanes (sm_100 FFMA2 / FMUL2 operands)
__device__ __forceinline__ jac_f2 jac_pack(float lo, float hi) {
    jac_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float jac_hsum(jac_f2 v) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    return lo + hi;
}
__device__ __forceinline__ jac_f2 jac_fma2(jac_f2 a, jac_f2 b, jac_f2 c) {
    jac_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ jac_f2 jac_mul2(jac_f2 a, jac_f2 b) {
    jac_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ jac_f2 jac_add2(jac_f2 a, jac_f2 b) {
    jac_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float jac_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// rsqrt with one Newton step: the rotation's cs^2 + sn^2 = 1 rests on it, and the 2-ulp MUFU result alone lets the
// column norms (= the eigenvalues) drift by a random walk of ~3e-7 per rotation over ~2000 rotations per column
__device__ __forceinline__ float jac_rsqrt_refined(float x) {
    const float r = rsqrtf(x);
    return r * fmaf(-0.5f * x, r * r, 1.5f);
}
// Quadratic convergence: a sweep whose largest pre-rotation cosine was below kJacobiSmallAngle leaves cosines of the
// order of its square (1e-9), far below the tolerance, so it is the last one - the sweep that only verifies that nothing
// rotates any more (10 % of the pooled eigen-solver) is skipped.
constexpr float kJacobiSmallAngle = 3.0e-5f;
// One rotation of the column pairs held by the four 8-lane groups of this warp.  ALL 32 lanes must call it together.
template <int CHUNKS>
__device__ __forceinline__ int jac_rotate_regs(ulonglong2 (&x)[CHUNKS], ulonglong2 (&y)[CHUNKS], float tol) {
    jac_f2 al2 = 0ull, be2 = 0ull, ga2 = 0ull, al2b = 0ull, be2b = 0ull, ga2b = 0ull;      // 0ull = (+0.f, +0.f)
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        al2 = jac_fma2(x[c].x, x[c].x, al2); al2b = jac_fma2(x[c].y, x[c].y, al2b);
        be2 = jac_fma2(y[c].x, y[c].x, be2); be2b = jac_fma2(y[c].y, y[c].y, be2b);
        ga2 = jac_fma2(x[c].x, y[c].x, ga2); ga2b = jac_fma2(x[c].y, y[c].y, ga2b);
    }
    float al = jac_hsum(jac_add2(al2, al2b)), be = jac_hsum(jac_add2(be2, be2b)), ga = jac_hsum(jac_add2(ga2, ga2b));
#pragma unroll
    for (int o = JAC_GROUP / 2; o > 0; o >>= 1) {
        al += __shfl_xor_sync(0xffffffffu, al, o);
        be += __shfl_xor_sync(0xffffffffu, be, o);
        ga += __shfl_xor_sync(0xffffffffu, ga, o);
    }
    // |theta| <= pi/4 with tan 2theta = h / d:  cos 2theta = |d| / r,  sin 2theta = sign(d) h / r,  r = hypot(d, h)
    const float d = be - al, h = ga + ga;
    const float r2 = fmaf(d, d, h * h);
    const float rinv = jac_rsqrt_refined(r2);
    const float y2 = fmaf(0.5f * fabsf(d), rinv, 0.5f);               // cos^2 theta, in [0.5, 1]
    const float icy = jac_rsqrt_refined(y2);
    const float sn_abs = 0.5f * h * rinv * icy;                       // sin 2theta / (2 cos theta), sign of h
    // (off the chain) already orthogonal, or d^2 + h^2 outside the fp32 range: leave the pair alone
    const float gn = jac_sqrt_approx(al * be);
    const bool rot = fabsf(ga) > tol * gn && r2 > 1e-36f && r2 < 1e36f;
    const bool big = fabsf(ga) > kJacobiSmallAngle * gn;              // a rotation that is not yet in the quadratic end game
    const float cs = rot ? y2 * icy : 1.f;
    const float sn = rot ? (d < 0.f ? -sn_abs : sn_abs) : 0.f;
    const jac_f2 cs2 = jac_pack(cs, cs), sn2 = jac_pack(sn, sn), nsn2 = jac_pack(-sn, -sn);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        ulonglong2 xn, yn;
        xn.x = jac_fma2(cs2, x[c].x, jac_mul2(nsn2, y[c].x)); yn.x = jac_fma2(sn2, x[c].x, jac_mul2(cs2, y[c].x));
        xn.y = jac_fma2(cs2, x[c].y, jac_mul2(nsn2, y[c].y)); yn.y = jac_fma2(sn2, x[c].y, jac_mul2(cs2, y[c].y));
        x[c] = xn; y[c] = yn;
    }
    return rot ? (big ? 3 : 1) : 0;                                   // bit 0: rotated, bit 1: by more than kJacobiSmallAngle
}
// generic-pointer column load / store (initial load from and final store to cluster rank 0)
template <int CHUNKS>
__device__ __forceinline__ void jac_ld(const float* __restrict__ col, int ld, int gl, bool valid, ulonglong2 (&v)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int r = c * JAC_CHUNK_ROWS + gl * 4;
        v[c] = (valid && r < ld) ? *reinterpret_cast<const ulonglong2*>(col + r) : make_ulonglong2(0ull, 0ull);
    }
}
template <int CHUNKS>
__device__ __forceinline__ void jac_st(float* __restrict__ col, int ld, int gl, bool valid, const ulonglong2 (&v)[CHUNKS]) {
    if (!valid) return;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int r = c * JAC_CHUNK_ROWS + gl * 4;
        if (r < ld) *reinterpret_cast<ulonglong2*>(col + r) = v[c];
    }
}
// the same on 32-bit shared-memory addresses of this CTA (the mailboxes; predicated, no branch around the warp)
template <int CHUNKS>
__device__ __forceinline__ void jac_lds(uint32_t col, int ld, int gl, bool valid, ulonglong2 (&v)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int r = c * JAC_CHUNK_ROWS + gl * 4;
        ulonglong2 t = make_ulonglong2(0ull, 0ull);
        if (valid && r < ld) asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(t.x), "=l"(t.y) : "r"(col + r * 4) : "memory");
        v[c] = t;
    }
}
template <int CHUNKS>
__device__ __forceinline__ void jac_sts(uint32_t col, int ld, int gl, bool valid, const ulonglong2 (&v)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int r = c * JAC_CHUNK_ROWS + gl * 4;
        if (valid && r < ld) asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(col + r * 4), "l"(v[c].x), "l"(v[c].y) : "memory");
    }
}
__device__ __forceinline__ void jac_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t jac_mapa(uint32_t saddr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// 16 bytes to the shared memory of another CTA; the receiver's mbarrier is credited with 16 bytes on arrival
__device__ __forceinline__ void jac_st_async(uint32_t raddr, ulonglong2 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "l"(v.x), "l"(v.y),
                 "r"(rbar)
                 : "memory");
}
template <int CHUNKS>
__device__ __forceinline__ void jac_push(uint32_t rcol, uint32_t rbar, int ld, int gl, bool valid, const ulonglong2 (&v)[CHUNKS]) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int r = c * JAC_CHUNK_ROWS + gl * 4;
        if (valid && r < ld) jac_st_async(rcol + r * 4, v[c], rbar);
    }
}
__device__ __forceinline__ void jac_bar_active(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
// OR over the threads of named barrier 1
__device__ __forceinline__ int jac_bar_active_or(int nthreads, int v) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.or.pred p, 1, %2, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(v), "r"(nthreads)
        : "memory");
    return static_cast<int>(r);
}

// ------------------------------------------------------------------------------------------------------------------
// Round-robin variant on ONE CTA, columns in shared memory (the small k x k problems of the angles kernel and any
// shape the cluster variant does not take).  n-1 steps per sweep, n/2 disjoint pairs per step, one 8-lane group per
// pair; only the warps that hold pairs take part (named barrier), the rest of the CTA waits at the final barrier
// instead of walking through ~(n-1) x sweeps full-CTA barriers (that was 21 % of the angles kernel's samples).
// Returns the number of sweeps executed (valid in thread 0).  All threads of the CTA must call it.
// n_cols may be odd (the virtual last column is skipped).  Rows [m, ld) of every column must be zero.
// ------------------------------------------------------------------------------------------------------------------
template <int CHUNKS>
__device__ int jacobi_orthogonalize(float* __restrict__ A, int ld, int n_cols, float tol, int max_sweeps) {
    const int n = (n_cols + 1) & ~1;
    const int half = n / 2;
    const int group = threadIdx.x / JAC_GROUP;
    const int gl = threadIdx.x % JAC_GROUP;
    const int n_groups = min(half, static_cast<int>(blockDim.x) / JAC_GROUP);
    const int n_active = min((n_groups * JAC_GROUP + 31) & ~31, static_cast<int>(blockDim.x));
    const int act_groups = n_active / JAC_GROUP;            // groups in the participating warps (a multiple of 4)
    const uint32_t a_s = smem_u32(A);
    int sweep = 0;
    if (n_cols >= 2 && static_cast<int>(threadIdx.x) < n_active) {
        for (; sweep < max_sweeps; ++sweep) {
            int rotated = 0;
            for (int s = 0; s < n - 1; ++s) {
                for (int kb = 0; kb < half; kb += act_groups) {         // same trip count for every lane of a warp
                    const int k = kb + group;
                    int p = 0, q = 0;
                    if (k < half) jacobi_pair(n, s, k, p, q);
                    const bool ok = k < half && p < n_cols && q < n_cols;   // else: no pair, or the virtual padding column
                    const uint32_t cp = a_s + static_cast<uint32_t>(ok ? p : 0) * ld * 4;
                    const uint32_t cq = a_s + static_cast<uint32_t>(ok ? q : 0) * ld * 4;
                    ulonglong2 x[CHUNKS], y[CHUNKS];
                    jac_lds<CHUNKS>(cp, ld, gl, ok, x);
                    jac_lds<CHUNKS>(cq, ld, gl, ok, y);
                    const int r = jac_rotate_regs<CHUNKS>(x, y, tol);
                    rotated |= r;
                    jac_sts<CHUNKS>(cp, ld, gl, ok && (r & 1), x);
                    jac_sts<CHUNKS>(cq, ld, gl, ok && (r & 1), y);
                }
                jac_bar_active(n_active);
            }
            if (!jac_bar_active_or(n_active, rotated & 2)) { ++sweep; break; }     // see kJacobiSmallAngle
        }
    }
    __syncthreads();
    return sweep;
}


// COMPACT = false: the matrix is in the shared memory of cluster rank 0 at `A` and every CTA uses its own copy of that
//   region as mailboxes (slot of group g = column 2g).  COMPACT = true: the matrix is in GLOBAL memory at `A` (problems
//   that do not fit one SM: n up to 384 with 12 chunks) and `mail` is this CTA's shared mailbox area of gpc x ld floats
//   (slot of a group = its index inside the CTA).
template <int CHUNKS, bool COMPACT = false>
__device__ int jacobi_orthogonalize_oddeven_cluster(float* __restrict__ A, int ld, int n_cols, float tol, int max_sweeps, float* inbox,
                                                    uint64_t* bars, int* flags, float* mail = nullptr) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int C = static_cast<int>(cluster.num_blocks()), crank = static_cast<int>(cluster.block_rank());
    const int n = (n_cols + 1) & ~1;
    const int m = n / 2;                                    // groups in the whole cluster
    const int gpc = (m + C - 1) / C;                        // groups per CTA
    const int n_active = (gpc * JAC_GROUP + 31) & ~31;      // threads of the warps that hold groups
    const int gloc = threadIdx.x / JAC_GROUP;
    const int gl = threadIdx.x % JAC_GROUP;
    const int g = crank * gpc + gloc;
    const bool active = gloc < gpc && g < m;
    const bool q_real = active && (2 * g + 1 < n_cols);     // the virtual last column of an odd n_cols is all zeros
    const bool has_next = active && g + 1 < m;              // an odd-step partner exists
    const bool next_remote = has_next && gloc == gpc - 1;   // ... and lives in CTA crank + 1
    const bool next_local = has_next && !next_remote;
    const bool prev_remote = active && gloc == 0 && crank > 0;   // this group is such a partner for CTA crank - 1
    uint64_t* bar_in = bars;                                // credited by group g+1's pushes into `inbox`
    uint64_t* bar_back = bars + 1;                          // credited by group g-1's pushes into this group's mailbox
    if (threadIdx.x == 0) {
        mbar_init(bar_in, 1);
        mbar_init(bar_back, 1);
        fence_mbar_init();
    }
    float* A0 = COMPACT ? A : cluster.map_shared_rank(A, 0);                             // the matrix itself
    const uint32_t a_s = smem_u32(COMPACT ? mail : A);
    const int slotP = COMPACT ? gloc : 2 * g, slotN = COMPACT ? gloc + 1 : 2 * g + 2, slotR = COMPACT ? 0 : 2 * g + 2;
    const uint32_t colP = a_s + static_cast<uint32_t>(active ? slotP : 0) * ld * 4;      // own mailbox = own even column slot
    const uint32_t colN = next_remote ? smem_u32(inbox) : a_s + static_cast<uint32_t>(next_local ? slotN : 0) * ld * 4;
    // addresses in the shared::cluster window of the neighbouring CTAs
    const uint32_t r_inbox = jac_mapa(smem_u32(inbox), prev_remote ? crank - 1 : crank);
    const uint32_t r_bar_in = jac_mapa(smem_u32(bar_in), prev_remote ? crank - 1 : crank);
    const uint32_t r_mail = jac_mapa(a_s + static_cast<uint32_t>(next_remote ? slotR : 0) * ld * 4, next_remote ? crank + 1 : crank);
    const uint32_t r_bar_back = jac_mapa(smem_u32(bar_back), next_remote ? crank + 1 : crank);
    const uint32_t col_bytes = static_cast<uint32_t>(ld) * 4u;
    ulonglong2 P[CHUNKS], Q[CHUNKS];
    jac_ld<CHUNKS>(A0 + static_cast<size_t>(active ? 2 * g : 0) * ld, ld, gl, active, P);
    jac_ld<CHUNKS>(A0 + static_cast<size_t>(q_real ? 2 * g + 1 : 0) * ld, ld, gl, q_real, Q);
    jac_cluster_sync();                                     // mbarriers initialised everywhere, initial loads done
    int sweep = 0;
    uint32_t phase = 0;
    if (n_cols >= 2) {
        for (; sweep < max_sweeps; ++sweep) {
            int rotated = 0;
            if (static_cast<int>(threadIdx.x) < n_active) {
                for (int pairstep = 0; pairstep < m; ++pairstep, phase ^= 1) {
                    if (gl == 0) {                          // arm this pair-step's deliveries
                        if (next_remote) mbar_arrive_expect_tx(bar_in, col_bytes);
                        if (prev_remote) mbar_arrive_expect_tx(bar_back, col_bytes);
                    }
                    // even step: positions (2g, 2g+1) = (P, Q); afterwards position 2g holds Q, position 2g+1 holds P
                    rotated |= jac_rotate_regs<CHUNKS>(P, Q, tol);
                    jac_push<CHUNKS>(r_inbox, r_bar_in, ld, gl, prev_remote, Q);          // to the last group of the previous CTA, or
                    jac_sts<CHUNKS>(colP, ld, gl, active && !prev_remote, Q);             // mail own position-2g column locally
                    jac_bar_active(n_active);
                    // odd step: positions (2g+1, 2g+2) = (P, even column of group g+1); groups without a partner rotate
                    // against a zero column (a no-op) to keep the warp converged
                    if (next_remote) mbar_wait(bar_in, phase);
                    __syncwarp();
                    jac_lds<CHUNKS>(colN, ld, gl, has_next, Q);
                    rotated |= jac_rotate_regs<CHUNKS>(P, Q, tol);
                    jac_push<CHUNKS>(r_mail, r_bar_back, ld, gl, next_remote, P);         // rotated old 2g+1 moves on to position 2g+2;
                    jac_sts<CHUNKS>(colN, ld, gl, next_local, P);                         // Q stays as position 2g+1
                    jac_bar_active(n_active);
                    if (prev_remote) mbar_wait(bar_back, phase);
                    __syncwarp();
                    if (!has_next) {                        // last group of all: its odd column (in P) was idle
#pragma unroll
                        for (int c = 0; c < CHUNKS; ++c) Q[c] = P[c];
                    }
                    jac_lds<CHUNKS>(colP, ld, gl, active, P);   // new position 2g from the mailbox; (P, Q) = (2g, 2g+1) again
                }
            }
            // did any CTA of the cluster rotate in this sweep?  (flags double-buffered by sweep parity)
            const int any_local = (__syncthreads_or(rotated & 1) ? 1 : 0) | (__syncthreads_or(rotated & 2) ? 2 : 0);
            int* fl = flags + (sweep & 1) * 8;
            if (threadIdx.x < C) cluster.map_shared_rank(fl, threadIdx.x)[crank] = any_local;
            jac_cluster_sync();
            int any = 0;
            for (int r = 0; r < C; ++r) any |= fl[r];
            if (!(any & 2)) { ++sweep; break; }             // nothing rotated, or only by end-game angles: converged
        }
    }
    // Every sweep reverses the order of the positions (n always-swap steps of the transposition network).  That only
    // matters for an odd n_cols: the virtual zero column starts at the last position and sits at position 0 after an odd
    // number of sweeps - the real columns are then positions 1 .. n-1 and go to columns 0 .. n_cols-1.
    const bool reversed = (n_cols & 1) && (sweep & 1);
    const int col_p = 2 * g - (reversed ? 1 : 0), col_q = 2 * g + 1 - (reversed ? 1 : 0);
    const bool st_p = active && col_p >= 0 && col_p < n_cols, st_q = active && col_q < n_cols;
    jac_st<CHUNKS>(A0 + static_cast<size_t>(st_p ? col_p : 0) * ld, ld, gl, st_p, P);
    jac_st<CHUNKS>(A0 + static_cast<size_t>(st_q ? col_q : 0) * ld, ld, gl, st_q, Q);
    jac_cluster_sync();
    return sweep;
}

// ------------------------------------------------------------------------------------------------------------------
// Round-robin variant on ONE CTA with the columns in GLOBAL memory (L2-resident), for matrices that do not fit one
// SM's shared memory (pooled eigenproblems of size 384, marchenko_pastur_rank on unprojected D_t-wide features) -
// any column length.  A pair of columns is owned by a GL-lane group (GL = 4 or 8) and streamed twice: once for the
// three inner products, once for the rotation.  Same rotation, tolerance and end-game exit as the variants above.
// All threads of the CTA must call it; returns the sweep count.  ld a multiple of 4, A 16-byte aligned.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 jac_ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 jac_rot4(float c, float s, const float4& x, const float4& y) {      // c x + s y
    return make_float4(fmaf(c, x.x, s * y.x), fmaf(c, x.y, s * y.y), fmaf(c, x.z, s * y.z), fmaf(c, x.w, s * y.w));
}
constexpr int JG_U = 6;                     // 16-byte loads per column in flight per lane
template <int GL>
__device__ int jacobi_orthogonalize_global(float* __restrict__ A, int ld, int n_cols, float tol, int max_sweeps) {
    const int n = (n_cols + 1) & ~1;
    const int half = n / 2;
    const int group = threadIdx.x / GL;
    const int gl = threadIdx.x % GL;
    const int n_groups = static_cast<int>(blockDim.x) / GL;
    int sweep = 0;
    if (n_cols >= 2) {
        for (; sweep < max_sweeps; ++sweep) {
            int rotated = 0;
            for (int s = 0; s < n - 1; ++s) {
                for (int kb = 0; kb < half; kb += n_groups) {           // same trip count for every thread
                    const int k = kb + group;
                    int p = 0, q = 0;
                    if (k < half) jacobi_pair(n, s, k, p, q);
                    const bool ok = k < half && p < n_cols && q < n_cols;
                    float* cp = A + static_cast<size_t>(ok ? p : 0) * ld;
                    float* cq = A + static_cast<size_t>(ok ? q : 0) * ld;
                    float al = 0.f, be = 0.f, ga = 0.f;
                    if (ok) {
                        int r = gl * 4;
                        for (; r + (JG_U - 1) * GL * 4 < ld; r += JG_U * GL * 4) {  // JG_U 16-byte loads of each column in flight
                            float4 x[JG_U], y[JG_U];
#pragma unroll
                            for (int u = 0; u < JG_U; ++u) { x[u] = jac_ldcg4(cp + r + u * GL * 4); y[u] = jac_ldcg4(cq + r + u * GL * 4); }
#pragma unroll
                            for (int u = 0; u < JG_U; ++u) {
                                al = fmaf(x[u].x, x[u].x, fmaf(x[u].y, x[u].y, fmaf(x[u].z, x[u].z, fmaf(x[u].w, x[u].w, al))));
                                be = fmaf(y[u].x, y[u].x, fmaf(y[u].y, y[u].y, fmaf(y[u].z, y[u].z, fmaf(y[u].w, y[u].w, be))));
                                ga = fmaf(x[u].x, y[u].x, fmaf(x[u].y, y[u].y, fmaf(x[u].z, y[u].z, fmaf(x[u].w, y[u].w, ga))));
                            }
                        }
                        for (; r < ld; r += GL * 4) {
                            const float4 x = jac_ldcg4(cp + r), y = jac_ldcg4(cq + r);
                            al = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, al))));
                            be = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, be))));
                            ga = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, ga))));
                        }
                    }
#pragma unroll
                    for (int o = GL / 2; o > 0; o >>= 1) {
                        al += __shfl_xor_sync(0xffffffffu, al, o);
                        be += __shfl_xor_sync(0xffffffffu, be, o);
                        ga += __shfl_xor_sync(0xffffffffu, ga, o);
                    }
                    const float d = be - al, h = ga + ga;
                    const float r2 = fmaf(d, d, h * h);
                    const float rinv = jac_rsqrt_refined(r2);
                    const float y2 = fmaf(0.5f * fabsf(d), rinv, 0.5f);
                    const float icy = jac_rsqrt_refined(y2);
                    const float sn_abs = 0.5f * h * rinv * icy;
                    const float gn = jac_sqrt_approx(al * be);
                    const bool rot = ok && fabsf(ga) > tol * gn && r2 > 1e-36f && r2 < 1e36f;
                    if (rot) {
                        rotated |= fabsf(ga) > kJacobiSmallAngle * gn ? 3 : 1;
                        const float cs = y2 * icy, sn = d < 0.f ? -sn_abs : sn_abs;
                        int r = gl * 4;
                        for (; r + (JG_U - 1) * GL * 4 < ld; r += JG_U * GL * 4) {
                            float4 x[JG_U], y[JG_U];
#pragma unroll
                            for (int u = 0; u < JG_U; ++u) { x[u] = jac_ldcg4(cp + r + u * GL * 4); y[u] = jac_ldcg4(cq + r + u * GL * 4); }
#pragma unroll
                            for (int u = 0; u < JG_U; ++u) {
                                *reinterpret_cast<float4*>(cp + r + u * GL * 4) = jac_rot4(cs, -sn, x[u], y[u]);
                                *reinterpret_cast<float4*>(cq + r + u * GL * 4) = jac_rot4(sn, cs, x[u], y[u]);
                            }
                        }
                        for (; r < ld; r += GL * 4) {
                            const float4 x = jac_ldcg4(cp + r), y = jac_ldcg4(cq + r);
                            *reinterpret_cast<float4*>(cp + r) = jac_rot4(cs, -sn, x, y);
                            *reinterpret_cast<float4*>(cq + r) = jac_rot4(sn, cs, x, y);
                        }
                    }
                }
                __syncthreads();
            }
            if (!__syncthreads_or(rotated & 2)) { ++sweep; break; }     // see kJacobiSmallAngle
        }
    }
    __syncthreads();
    return sweep;
}

// Column norms (sqrt of sum of squares) of the first n_cols columns -> out[n_cols].  One warp per column.
__device__ inline void column_norms(const float* __restrict__ A, int ld, int m, int n_cols, float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = warp; c < n_cols; c += nw) {
        const float* col = A + static_cast<size_t>(c) * ld;
        float s = 0.f;
        for (int r = lane; r < m; r += 32) s = fmaf(col[r], col[r], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[c] = sqrtf(s);
    }
}

// order[r] = index of the column with the r-th largest value (descending; ties by index).  n <= blockDim-strided.
__device__ inline void rank_descending(const float* __restrict__ val, int n, int* __restrict__ order) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = val[i];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const float u = val[j];
            r += (u > v) || (u == v && j < i);
        }
        order[r] = i;
    }
}

}  // namespace basd
