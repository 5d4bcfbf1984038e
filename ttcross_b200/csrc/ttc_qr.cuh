// =============================================================================
// ttc_qr.cuh — thin Householder QR of a tall-skinny block: ort0_d of the reference
// (lib/ort.f90:17-81 = LAPACK dgeqrf + dorgqr), the kernel named by SURVEY §8 row a21.
//
// Shapes on this path are (r*n) x r unfoldings of TT cores: 8224 x 32 (config B), 12336 x 48 (C),
// 32832 x 64 (D).  The block is dealt out by contiguous row chunks to the CTAs of ONE cooperative
// launch and stays in shared memory for the whole factorisation (148 CTAs x <= 113 KB at D); only
// the per-column reductions — the sum of squares for dlarfg and the products v^T A(:,j) for dlarf —
// cross CTAs, through small partial arrays in HBM and two grid barriers per column.  Reflectors
// follow LAPACK's unblocked dgeqr2/dlarfg/dlarf/dorg2r conventions (beta = -sign(alpha)*norm, v(1) = 1),
// so R carries LAPACK's signs and Q is the explicit m x n factor.  Summation order inside the
// reductions differs from any particular LAPACK build (the reference links an unpinned one), hence
// parity is to rounding, not bit-exact.  Compiled with -fmad=false like the rest of the library.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>

namespace ttc {
namespace cgq = cooperative_groups;

constexpr int QR_THREADS = 256;

// block-wide sum in a fixed order (warp shuffles, then warp 0); result in every thread
__device__ __forceinline__ double qr_block_sum(double x, double* sh /*[33]*/) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = x;
    __syncthreads();
    if (w == 0) {
        double y = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) y += __shfl_down_sync(0xffffffffu, y, o);
        if (lane == 0) sh[32] = y;
    }
    __syncthreads();
    return sh[32];
}

// a: m x n column-major (lda); q: m x n out (ldq = m); r: n x n out (upper triangular, zeros below)
// part[gridDim.x], partw[2 * gridDim.x * n] (double-buffered), head[n + 2]: scratch in HBM.  rpb rows per CTA.  dynamic smem: rpb*n + n doubles.
__global__ void __launch_bounds__(QR_THREADS) k_qr_panel(const double* __restrict__ a, int m, int n, int lda, double* __restrict__ q,
                                                       double* __restrict__ r, double* part, double* partw, double* head, int rpb) {
    cgq::grid_group grid = cgq::this_grid();
    extern __shared__ double smem[];
    __shared__ double shr[33];
    double* S = smem;                         // S[i + rows*j], i local row
    double* wv = smem + (size_t)rpb * n;      // wv[j]: reduced v^T A(:,j) of the current reflector
    const int row0 = blockIdx.x * rpb;
    const int rows = max(0, min(rpb, m - row0));
    const int G = gridDim.x;
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; S[e] = a[(size_t)(row0 + i) + (size_t)lda * j]; }
    __syncthreads();
    double tau_k = 0.0;

    // ---------------- dgeqr2: reflectors H(1) ... H(n)
    for (int k = 0; k < n; ++k) {
        // (a) partial sum of squares below the diagonal; the owner of row k publishes alpha
        double ss = 0.0;
        for (int i = threadIdx.x; i < rows; i += blockDim.x) if (row0 + i > k) { const double x = S[i + rows * k]; ss += x * x; }
        ss = qr_block_sum(ss, shr);
        if (threadIdx.x == 0) {
            part[blockIdx.x] = ss;
            if (k >= row0 && k < row0 + rows) head[0] = S[(k - row0) + rows * k];
        }
        grid.sync();
        // (b) dlarfg, identically in every CTA
        double xn2 = 0.0;
        for (int b = 0; b < G; ++b) xn2 += part[b];
        const double alpha = head[0];
        double beta = alpha, scale = 0.0;
        tau_k = 0.0;
        if (xn2 != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
            tau_k = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        // (c) v = x * scale (v(k) = 1 implicit), partial w_j = v^T A(:, j), j > k
        for (int i = threadIdx.x; i < rows; i += blockDim.x) if (row0 + i > k) S[i + rows * k] *= scale;
        __syncthreads();
        for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x) {
            double w = 0.0;
            for (int i = 0; i < rows; ++i) {
                const int gi = row0 + i;
                if (gi < k) continue;
                const double vi = (gi == k) ? 1.0 : S[i + rows * k];
                w += vi * S[i + rows * j];
            }
            partw[(size_t)blockIdx.x * n + j] = w;
        }
        grid.sync();
        // (d) A(k:m, j) -= tau * w_j * v
        for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x) { double w = 0.0; for (int b = 0; b < G; ++b) w += partw[(size_t)b * n + j]; wv[j] = w; }
        if (threadIdx.x == 0) wv[k] = tau_k;
        __syncthreads();
        for (int e = threadIdx.x; e < rows * (n - k - 1); e += blockDim.x) {
            const int jj = e / rows, i = e - jj * rows, j = k + 1 + jj, gi = row0 + i;
            if (gi < k) continue;
            const double vi = (gi == k) ? 1.0 : S[i + rows * k];
            S[i + rows * j] -= tau_k * wv[j] * vi;
        }
        if (k >= row0 && k < row0 + rows && threadIdx.x == 0) { S[(k - row0) + rows * k] = beta; }
        // tau(k) is needed again by dorg2r: keep it in the (otherwise unused) strictly-lower part of r
        if (blockIdx.x == 0 && threadIdx.x == 0) head[2 + k] = tau_k;
        __syncthreads();
    }
    // R = upper triangle of the first n rows
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) {
        const int j = e / rows, i = e - j * rows, gi = row0 + i;
        if (gi < n) r[gi + (size_t)n * j] = (gi <= j) ? S[e] : 0.0;
    }
    grid.sync();

    // ---------------- dorg2r: Q = H(1) ... H(n) applied to the first n columns of the identity
    for (int k = n - 1; k >= 0; --k) {
        const double tk = head[2 + k];
        if (k < n - 1) {
            double* pw = partw + (size_t)(k & 1) * G * n;      // double-buffered: one grid barrier per reflector
            // w_j = v^T Q(k:m, j) for j > k (v(k) = 1)
            for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x) {
                double w = 0.0;
                for (int i = 0; i < rows; ++i) {
                    const int gi = row0 + i;
                    if (gi < k) continue;
                    const double vi = (gi == k) ? 1.0 : S[i + rows * k];
                    w += vi * S[i + rows * j];
                }
                pw[(size_t)blockIdx.x * n + j] = w;
            }
            grid.sync();
            for (int j = k + 1 + threadIdx.x; j < n; j += blockDim.x) { double w = 0.0; for (int b = 0; b < G; ++b) w += pw[(size_t)b * n + j]; wv[j] = w; }
            __syncthreads();
            for (int e = threadIdx.x; e < rows * (n - k - 1); e += blockDim.x) {
                const int jj = e / rows, i = e - jj * rows, j = k + 1 + jj, gi = row0 + i;
                if (gi < k) continue;
                const double vi = (gi == k) ? 1.0 : S[i + rows * k];
                S[i + rows * j] -= tk * wv[j] * vi;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < rows; i += blockDim.x) {
            const int gi = row0 + i;
            double x = S[i + rows * k];
            if (gi > k) x = -tk * x; else if (gi == k) x = 1.0 - tk; else x = 0.0;
            S[i + rows * k] = x;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; q[(size_t)(row0 + i) + (size_t)m * j] = S[e]; }
}

// =============================================================================
// TSQR (round 2): communication-avoiding QR for the tall-skinny case (m >= 4n, n <= 64) with LAPACK's result.
//
// k_qr_panel above runs the unblocked algorithm over the whole grid: two grid barriers per column (64 columns x 2 phases).
// Here the rows are dealt out in blocks of ~4n; every block is factored INSIDE one CTA (shared memory, block barriers only),
// the n x n R factors are stacked four at a time and factored again, level by level, until one R is left
// (k_tsqr_factor, one launch per level); then every first-level block multiplies its explicit Q by the chain of n x n
// slices of the upper levels' Q factors (k_tsqr_apply: Q = Q1 Q2 ... QL, a dense (rows x n)(n x n) product per block).
//
// A QR is unique up to the signs of R's rows.  LAPACK's signs (beta = -sign(alpha) |x| at every step of the UNBLOCKED
// algorithm) depend on the history of that algorithm, but they can be recovered from any QR by "Householder
// reconstruction" (Ballard, Demmel, Grigori, Jacquelin, Nguyen, Solomonik 2014): run an LU factorisation without pivoting on
// the top n x n block of Q, choosing S_kk = -sgn(current diagonal entry) and subtracting S_kk from it before eliminating;
// then R_H = S R and Q_H = Q S are the Householder (LAPACK) factors.  k_tsqr_sign does that in one CTA, k_tsqr_scale applies
// S to the columns of Q.  Checked against numpy.linalg.qr (tests/test_qr.py, unchanged).
// =============================================================================
constexpr int TSQR_THREADS = 1024;     // k_tsqr_apply / k_tsqr_sign
// ---- block factorisation: columns in REGISTERS, reflectors published through shared memory, no block barrier per column.
// A CTA of 16 warps factors a block of rows <= 256, n <= 64.  Warp w owns columns w, w+16, w+32, w+48; a lane holds rows
// lane, lane+32, ... of each (8 per column).  The algorithm is left-looking per warp: reflector k is generated by the owner of
// column k as soon as H(0..k-1) have been applied to THAT column, written to V (shared memory) and announced through a
// counter; every other warp applies H(k) to its columns when it sees the counter pass k.  The critical path of a step is one
// warp's apply (dot, 5 shuffles, update) + generate (sum of squares, 5 shuffles, sqrt, two divisions) + the flag, ~0.45 us,
// instead of three 1024-thread barriers around shared-memory passes (1.35 us); the other warps' updates run beside it.
// dorg2r needs no synchronisation at all: column j of Q is H(0) ... H(j) e_j, a chain private to the warp that owns column j.
constexpr int TQ_THREADS = 512, TQ_WARPS = 16, TQ_RPL = 8, TQ_LDV = 32 * TQ_RPL;
constexpr unsigned TQ_FULL = 0xffffffffu;
// One mbarrier per reflector (arrival count 1): the owner arrives (release) once the reflector is in shared memory, the
// consumers wait on phase 0 with mbarrier.try_wait (acquire), which parks the warp in hardware instead of polling shared memory
// -- 15 warps spinning on a flag word saturate the shared-memory pipe the critical warp's shuffles go through (measured: a
// polled counter made the kernel 10x slower than the barrier-per-column version it replaced).
__device__ __forceinline__ unsigned tq_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tq_mbar_init(unsigned long long* bar) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tq_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tq_mbar_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tq_smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tq_mbar_wait(unsigned long long* bar) {
    unsigned ok = 0;
    for (int spins = 0; !ok && spins < (1 << 20); ++spins)          // (bounded: a lost reflector must not hang the GPU)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(tq_smem_u32(bar)) : "memory");
}

// H(k) = I - tau v v^T applied to column slots QA .. QB-1 of this warp (compile-time range: straight-line code, no predicates --
// the first version of this kernel spent half of its issue slots on FSEL / ISETP / BSSY around per-slice conditions,
// profiles/r02f_ncu_tsqr.txt).  vk: reflector k in shared memory, all TQ_LDV rows valid (zeros above row k and in the padding).
template <int NC, int QA, int QB>
__device__ __forceinline__ void tq_apply(double (&c)[NC][TQ_RPL], const double* vk, double tau, int lane) {
    if constexpr (QA < QB) {
    double v[TQ_RPL];
#pragma unroll
    for (int t = 0; t < TQ_RPL; ++t) v[t] = vk[lane + 32 * t];
    double w[NC];
#pragma unroll
    for (int q = QA; q < QB; ++q) {
        // (explicit FMA: this QR is held to LAPACK by tolerance, not bit for bit -- the library's -fmad=false is for the sweep)
        double wa = v[0] * c[q][0], wb = v[1] * c[q][1];
#pragma unroll
        for (int t = 2; t < TQ_RPL; t += 2) { wa = fma(v[t], c[q][t], wa); wb = fma(v[t + 1], c[q][t + 1], wb); }
        w[q] = wa + wb;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = QA; q < QB; ++q) w[q] += __shfl_xor_sync(TQ_FULL, w[q], o);
    }
#pragma unroll
    for (int q = QA; q < QB; ++q) {
        const double tw = -(tau * w[q]);
#pragma unroll
        for (int t = 0; t < TQ_RPL; ++t) c[q][t] = fma(tw, v[t], c[q][t]);
    }
    }
}
// the same for the runtime (warp-uniform) range q0 .. NC-1
template <int NC>
__device__ __forceinline__ void tq_apply_from(double (&c)[NC][TQ_RPL], int q0, const double* vk, double tau, int lane) {
    switch (q0) {
        case 0: tq_apply<NC, 0, NC>(c, vk, tau, lane); break;
        case 1: tq_apply<NC, (1 < NC ? 1 : NC), NC>(c, vk, tau, lane); break;
        case 2: tq_apply<NC, (2 < NC ? 2 : NC), NC>(c, vk, tau, lane); break;
        case 3: tq_apply<NC, (3 < NC ? 3 : NC), NC>(c, vk, tau, lane); break;
        default: break;
    }
}
template <int NC>
__device__ __forceinline__ void tq_apply_one(double (&c)[NC][TQ_RPL], int q, const double* vk, double tau, int lane) {
    switch (q) {
        case 0: tq_apply<NC, 0, 1>(c, vk, tau, lane); break;
        case 1: tq_apply<NC, (1 < NC ? 1 : NC), (1 < NC ? 2 : NC)>(c, vk, tau, lane); break;
        case 2: tq_apply<NC, (2 < NC ? 2 : NC), (2 < NC ? 3 : NC)>(c, vk, tau, lane); break;
        case 3: tq_apply<NC, (3 < NC ? 3 : NC), (3 < NC ? 4 : NC)>(c, vk, tau, lane); break;
        default: break;
    }
}
// dlarfg on one column (= global column k, every earlier reflector applied): beta on the diagonal, v below it (also left in the
// registers), published as reflector k with explicit zeros above row k
__device__ __forceinline__ double tq_generate(double (&cq)[TQ_RPL], int k, double* vk, double* tau_s, unsigned long long* bars, int lane) {
    double ss = 0.0, al = 0.0;
#pragma unroll
    for (int t = 0; t < TQ_RPL; ++t) {
        const int i = lane + 32 * t;
        const double x = cq[t];
        if (i > k) ss = fma(x, x, ss);
        if (i == k) al = x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(TQ_FULL, ss, o);
    const double alpha = __shfl_sync(TQ_FULL, al, k & 31);
    double beta = alpha, scale = 0.0, tk = 0.0;
    if (ss != 0.0) {
        beta = -copysign(sqrt(alpha * alpha + ss), alpha);
        tk = (beta - alpha) / beta;
        scale = 1.0 / (alpha - beta);
    }
#pragma unroll
    for (int t = 0; t < TQ_RPL; ++t) {
        const int i = lane + 32 * t;
        double x = 0.0;
        if (i > k) { x = cq[t] * scale; cq[t] = x; }
        else if (i == k) { cq[t] = beta; x = 1.0; }
        vk[i] = x;
    }
    if (lane == 0) tau_s[k] = tk;
    __threadfence_block();
    __syncwarp();
    if (lane == 0) tq_mbar_arrive(bars + k);
    return tk;
}
// the panel of one warp (compile-time recursion over its column slots)
template <int NC, int Q>
__device__ __forceinline__ void tq_panel(double (&c)[NC][TQ_RPL], int j0, int n, double* V, double* tau_s, unsigned long long* bars, int lane) {
    if constexpr (Q < NC) {
        const int j = j0 + Q;
        if (j < n) {
            const double tj = tq_generate(c[Q], j, V + (size_t)j * TQ_LDV, tau_s, bars, lane);
            tq_apply<NC, Q + 1, NC>(c, V + (size_t)j * TQ_LDV, tj, lane);
            tq_panel<NC, Q + 1>(c, j0, n, V, tau_s, bars, lane);
        }
    }
}
// One level of the tree, FACTOR phase only.  Node b factors `cnt` stacked source blocks:
//   level 1 : src = A (m x n, lda), block b = rows [b*m/G, (b+1)*m/G);                     raw block -> qout (ldq = m) at those rows
//   level>1 : src = the previous level's R factors (n x n each, contiguous), node b stacks R[F b .. F b + F-1];  raw block (cnt*n x n) -> qout + b*F n*n (ld F n)
// "raw block" = LAPACK's dgeqr2 storage (R on and above the diagonal, reflector vectors below), tau_out + b*n its scalars:
// forming the explicit Q from it (dorg2r) does not feed the next level -- only R does (rout + b*n*n, upper triangular, zeros
// below) -- so it is left to k_tsqr_formq / k_tsqr_leaf, which run once for all levels after the chain of factor launches.
// NC = columns per warp = ceil(n / 16).  dynamic smem: (n * TQ_LDV + 2 n) doubles.
template <int NC>
__global__ void __launch_bounds__(TQ_THREADS, 1) k_tsqr_factor(const double* __restrict__ src, int level, int m, int n, int lda, int G, int nsrc, int F,
                                                              double* __restrict__ rout, double* __restrict__ qout, int ldq, double* __restrict__ tau_out) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int rows, row0 = 0;
    if (level == 1) { row0 = (int)((long long)b * m / G); rows = (int)((long long)(b + 1) * m / G) - row0; }
    else { const int c0 = F * b; rows = min(F, nsrc - c0) * n; }
    double* V = smem; double* tau_s = V + (size_t)n * TQ_LDV;
    unsigned long long* bars = (unsigned long long*)(tau_s + n);
    for (int k = threadIdx.x; k < n; k += blockDim.x) tq_mbar_init(bars + k);
    // PANEL ownership in this phase: warp w holds the NC consecutive columns NC w .. NC w + NC-1, so the chain of dependent
    // reflectors changes warps only once per panel (16 hand-offs instead of 64) and stays inside one warp's registers between
    const int j0 = NC * wid;
    double c[NC][TQ_RPL];
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int j = j0 + q;
#pragma unroll
        for (int t = 0; t < TQ_RPL; ++t) {
            const int i = lane + 32 * t;
            double x = 0.0;
            if (j < n && i < rows) {
                if (level == 1) x = src[(size_t)(row0 + i) + (size_t)lda * j];
                else { const int blk = i / n, ii = i - blk * n; x = src[((size_t)(F * b + blk) * n + j) * n + ii]; }
            }
            c[q][t] = x;
        }
    }
    __syncthreads();
    // ---------------- dgeqr2
    if (j0 < n) {
        for (int k = 0; k < j0; ++k) {                     // reflectors of the earlier panels, as they are published
            tq_mbar_wait(bars + k);
            tq_apply<NC, 0, NC>(c, V + (size_t)k * TQ_LDV, tau_s[k], lane);
        }
        tq_panel<NC, 0>(c, j0, n, V, tau_s, bars, lane);   // my panel: generate, apply to the rest of the panel, next column
    }
    // ---------------- R and the raw block out
    double* qb = (level == 1) ? qout + row0 : qout + (size_t)b * F * n * n;
    const size_t ldb = (level == 1) ? (size_t)ldq : (size_t)F * n;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int j = j0 + q;
        if (j < n) {
#pragma unroll
            for (int t = 0; t < TQ_RPL; ++t) {
                const int i = lane + 32 * t;
                if (i < n) rout[(size_t)b * n * n + i + (size_t)n * j] = (i <= j) ? c[q][t] : 0.0;
                if (i < rows) qb[i + ldb * j] = c[q][t];
            }
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) tau_out[(size_t)b * n + k] = tau_s[k];
}
// dorg2r of a raw block held column-wise in registers: reload (raw block at qb, leading dimension ldb), publish the reflectors
// in shared memory, then every warp runs the chains of its own columns -- column j of Q is H(0) ... H(j) e_j -- with no
// synchronisation at all.  On return c holds the explicit Q block.
template <int NC>
__device__ __forceinline__ void tq_formq(double (&c)[NC][TQ_RPL], const double* __restrict__ qb, size_t ldb, int rows, int n,
                                         const double* __restrict__ tau_g, double* V, double* tau_s, int lane, int wid) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) tau_s[k] = tau_g[k];
    int jmax = -1;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int j = wid + TQ_WARPS * q;
#pragma unroll
        for (int t = 0; t < TQ_RPL; ++t) {
            const int i = lane + 32 * t;
            const double x = (j < n && i < rows && i > j) ? qb[i + ldb * j] : 0.0;
            c[q][t] = x;
            if (j < n) V[(size_t)j * TQ_LDV + i] = (i == j) ? 1.0 : x;
        }
        if (j < n) jmax = j;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int j = wid + TQ_WARPS * q;
        if (j < n) {
            const double tj = tau_s[j];
#pragma unroll
            for (int t = 0; t < TQ_RPL; ++t) { const int i = lane + 32 * t; c[q][t] = (i == j) ? 1.0 - tj : -tj * c[q][t]; }
        }
    }
    for (int k = jmax - 1; k >= 0; --k) {
        const int q0 = (k >= wid) ? ((k - wid) >> 4) + 1 : 0;
        tq_apply_from<NC>(c, q0, V + (size_t)k * TQ_LDV, tau_s[k], lane);
    }
}
// Explicit Q of every UPPER node (levels 2 .. L) in one launch, in place over the raw blocks.  blockIdx.x runs over the nodes
// of all upper levels: level l (0-based among the upper ones) owns blocks first[l] .. first[l+1]-1.
struct TsqrLevels { double* q[8]; const double* tau[8]; int first[9]; int nsrc[8]; int count; };
template <int NC>
__global__ void __launch_bounds__(TQ_THREADS, 1) k_tsqr_formq(int n, int F, TsqrLevels LV) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int l = 0;
    while (l + 1 < LV.count && (int)blockIdx.x >= LV.first[l + 1]) ++l;
    const int b = blockIdx.x - LV.first[l];
    const int rows = min(F, LV.nsrc[l] - F * b) * n;
    double* V = smem; double* tau_s = V + (size_t)n * TQ_LDV;
    double* qb = LV.q[l] + (size_t)b * F * n * n;
    const size_t ldb = (size_t)F * n;
    double c[NC][TQ_RPL];
    tq_formq<NC>(c, qb, ldb, rows, n, LV.tau[l] + (size_t)b * n, V, tau_s, lane, wid);
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const int j = wid + TQ_WARPS * q;
        if (j < n) {
#pragma unroll
            for (int t = 0; t < TQ_RPL; ++t) { const int i = lane + 32 * t; if (i < rows) qb[i + ldb * j] = c[q][t]; }
        }
    }
}
// First-level block b: explicit Q1 from its raw block (registers -> shared memory, over the reflector storage, which is dead by
// then), the chain M = slice(level 2) * slice(level 3) * ... formed in shared memory (each slice staged with coalesced loads
// first), and the final rows Q1 * M on 4 x 8 register tiles (rows ti + 64 r, columns tj + 8 c: conflict-free loads of Q1,
// broadcast loads of M; 12 shared-memory loads per 32 multiply-adds).  One launch replaces dorg2r + the block product and
// Q1 never travels through HBM.  dynamic smem: (n * TQ_LDV + n + 2 n^2) doubles.
template <int NC>
__global__ void __launch_bounds__(TQ_THREADS, 1) k_tsqr_leaf(double* __restrict__ q, int m, int n, int ldq, int G, int F, const double* __restrict__ tau1, TsqrLevels LV) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int row0 = (int)((long long)b * m / G), rows = (int)((long long)(b + 1) * m / G) - row0;
    double* V = smem; double* tau_s = V + (size_t)n * TQ_LDV; double* M = tau_s + n; double* T = M + n * n;
    double* qb = q + row0;
    double c[NC][TQ_RPL];
    tq_formq<NC>(c, qb, (size_t)ldq, rows, n, tau1 + (size_t)b * n, V, tau_s, lane, wid);
    if (LV.count == 0) {                                  // a single block: Q1 is the answer
#pragma unroll
        for (int q2 = 0; q2 < NC; ++q2) {
            const int j = wid + TQ_WARPS * q2;
            if (j < n) {
#pragma unroll
                for (int t = 0; t < TQ_RPL; ++t) { const int i = lane + 32 * t; if (i < rows) qb[i + (size_t)ldq * j] = c[q2][t]; }
            }
        }
        return;
    }
    __syncthreads();                                      // every chain has finished reading the reflectors
    double* QB = V;                                       // Q1 (TQ_LDV x n, leading dimension TQ_LDV)
#pragma unroll
    for (int q2 = 0; q2 < NC; ++q2) {
        const int j = wid + TQ_WARPS * q2;
        if (j < n) {
#pragma unroll
            for (int t = 0; t < TQ_RPL; ++t) QB[(size_t)j * TQ_LDV + lane + 32 * t] = c[q2][t];
        }
    }
    int idx = b;
    for (int l = 0; l < LV.count; ++l) {
        const int node = idx / F, slot = idx - F * node;
        const double* ql = LV.q[l] + (size_t)node * F * n * n + (size_t)slot * n;        // rows slot*n .. of the F n x n block
        double* dst = (l == 0) ? M : T;
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int j = e / n, i = e - j * n; dst[e] = ql[i + (size_t)F * n * j]; }
        __syncthreads();
        if (l > 0) {
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = threadIdx.x + u * TQ_THREADS;
                acc[u] = 0.0;
                if (e < n * n) {
                    const int j = e / n, i = e - j * n;
                    double t = 0.0;
                    for (int x = 0; x < n; ++x) t = fma(M[i + n * x], T[x + n * j], t);     // (a plain product of orthogonal factors: FMA is welcome)
                    acc[u] = t;
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int e = threadIdx.x + u * TQ_THREADS; if (e < n * n) M[e] = acc[u]; }
            __syncthreads();
        }
        idx = node;
    }
    const int ti = threadIdx.x & 63, tj = threadIdx.x >> 6;
    double acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) acc[r][cc] = 0.0;
    for (int x = 0; x < n; ++x) {
        double av[4], bv[8];
#pragma unroll
        for (int r = 0; r < 4; ++r) av[r] = QB[(size_t)x * TQ_LDV + ti + 64 * r];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) bv[cc] = M[x + n * min(tj + 8 * cc, n - 1)];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) acc[r][cc] = fma(av[r], bv[cc], acc[r][cc]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const int i = ti + 64 * r, j = tj + 8 * cc;
            if (i < rows && j < n) q[(size_t)(row0 + i) + (size_t)ldq * j] = acc[r][cc];
        }
}
// Householder reconstruction of LAPACK's signs: modified LU (no pivoting) of the top n x n block of Q; sgn[k] = S_kk; r <- S r.
__global__ void __launch_bounds__(TSQR_THREADS) k_tsqr_sign(const double* __restrict__ q, int n, int ldq, double* __restrict__ r, double* __restrict__ sgn) {
    extern __shared__ double smem[];
    double* W = smem; double* S = W + n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int j = e / n, i = e - j * n; W[e] = q[(size_t)i + (size_t)ldq * j]; }
    __syncthreads();
    // one barrier per step: every thread reads the (final) diagonal entry, forms the sign and the shifted pivot itself, and
    // divides on the fly (|pivot| >= 1 by construction); the multipliers are not needed afterwards, only the signs
    for (int k = 0; k < n; ++k) {
        const double d = W[k + n * k];
        const double sk = (d >= 0.0) ? -1.0 : 1.0;
        const double piv = d - sk;
        if (threadIdx.x == 0) S[k] = sk;
        const int i = threadIdx.x & 63, tj = threadIdx.x >> 6;          // thread = (row i, columns tj + 16 c): no index arithmetic in the chain
        if (i > k && i < n) {
            const double l = W[i + n * k] / piv;
#pragma unroll
            for (int c = 0; c < 4; ++c) { const int j = tj + 16 * c; if (j > k && j < n) W[i + n * j] -= l * W[k + n * j]; }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int i = e % n; r[e] = S[i] * r[e]; }
    for (int k = threadIdx.x; k < n; k += blockDim.x) sgn[k] = S[k];
}
__global__ void k_tsqr_scale(double* __restrict__ q, int m, int n, int ldq, const double* __restrict__ sgn) {
    const long long tot = (long long)m * n;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (long long)gridDim.x * blockDim.x) {
        const long long j = x / m; const long long i = x - j * m;
        if (sgn[j] < 0.0) q[i + (long long)ldq * j] = -q[i + (long long)ldq * j];
    }
}

// ----------------------------------------------------------------------------
// Support kernels of dtt_ort (lib/tt.f90:130-198): left-to-right orthogonalisation of the train.  Per core k:
//   QR of the (r(k-1) n(k)) x r(k) unfolding (k_qr_panel) -> R / ||R||_F, lognrm += log ||R||_F  (k_ort_rnorm)
//   core k <- Q (k_ort_store_q);  core k+1 <- R * core k+1 (k_ort_apply_r, dgemm 'n','n' order, beta = 0)
// and at the end the last core is normalised and every core is scaled by exp(lognrm / d) (k_ort_finish).
// Cores live in the sweep's padded layout: element (i,j,s) at i + ld*(j + n*s).
// ----------------------------------------------------------------------------
// acc[0] = running lognrm, acc[1] = norm of this R (for the record)
__global__ void k_ort_rnorm(double* r, int n, double* acc) {
    __shared__ double sh[33];
    double ss = 0.0;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) ss += r[e] * r[e];
    const double nrm = sqrt(qr_block_sum(ss, sh));
    if (nrm != 0.0) {
        const double sc = 1.0 / nrm;                     // dscal(mn*nn, 1.d0/nrm, mat, 1)
        for (int e = threadIdx.x; e < n * n; e += blockDim.x) r[e] = sc * r[e];
        if (threadIdx.x == 0) acc[0] = acc[0] + log(nrm);
    }
    if (threadIdx.x == 0) acc[1] = nrm;
}
// core(i + ld*(j + n*s)) <- q(e + mm*s), e = i + r0*j
__global__ void k_ort_store_q(const double* q, double* core, int r0, int n, int r1, int ld) {
    const long long mm = (long long)r0 * n, tot = mm * r1;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (long long)gridDim.x * blockDim.x) {
        const long long s = x / mm, e = x - s * mm;
        const int j = (int)(e / r0), i = (int)(e - (long long)j * r0);
        core[i + (long long)ld * (j + (long long)n * s)] = q[x];
    }
}
// next(i + ld*c) <- sum_l r(i + nn*l) * u(l + nn*c), l ascending from 0 (reference dgemm, beta = 0); u = packed copy of next
__global__ void k_ort_apply_r(const double* r, const double* u, double* next, int nn, long long kk, int ld) {
    const long long tot = (long long)nn * kk;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (long long)gridDim.x * blockDim.x) {
        const long long c = x / nn; const int i = (int)(x - c * nn);
        double t = 0.0;
        for (int l = 0; l < nn; ++l) t = t + u[l + (long long)nn * c] * r[i + (long long)nn * l];
        next[i + (long long)ld * c] = t;
    }
}
// norm of the (padded) last core into acc[2]; one CTA
__global__ void k_ort_lastnorm(const double* core, int r0, long long cols, int ld, double* acc) {
    __shared__ double sh[33];
    double ss = 0.0;
    const long long tot = (long long)r0 * cols;
    for (long long x = threadIdx.x; x < tot; x += blockDim.x) { const long long c = x / r0; const int i = (int)(x - c * r0); const double v = core[i + (long long)ld * c]; ss += v * v; }
    const double nrm = sqrt(qr_block_sum(ss, sh));
    if (threadIdx.x == 0) { acc[2] = nrm; if (nrm != 0.0) acc[0] = acc[0] + log(nrm); }
}
// scale one core by `pre` (1/||last|| for the last core, 1 otherwise) and by exp(acc[0] / d)
__global__ void k_ort_scale(double* core, int r0, long long cols, int ld, const double* acc, int d, int is_last) {
    const double nrm = exp(acc[0] / (double)d);
    const double pre = (is_last && acc[2] != 0.0) ? 1.0 / acc[2] : 1.0;
    const long long tot = (long long)r0 * cols;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (long long)gridDim.x * blockDim.x) {
        const long long c = x / r0; const int i = (int)(x - c * r0);
        double v = core[i + (long long)ld * c];
        if (is_last) v = pre * v;
        core[i + (long long)ld * c] = nrm * v;
    }
}

}  // namespace ttc
