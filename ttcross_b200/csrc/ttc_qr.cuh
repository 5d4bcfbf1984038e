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
constexpr int TSQR_THREADS = 1024;     // 32 warps: with two trailing columns per warp a reflector is applied in one round for n <= 64

// H(k) = I - tau v v^T (v(k) = 1, v(i) = S(i,k) below) applied to the trailing columns, TWO columns per warp at a time (they
// share the loads of v); dot product and update of a column stay inside its warp
__device__ __forceinline__ void cta_apply_reflector(double* S, int rows, int n, int k, double tk, double* next_ss = nullptr) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j0 = k + 1 + wid; j0 < n; j0 += 2 * nw) {
        const int j1 = j0 + nw;
        const bool two = j1 < n;
        const double* v = S + (size_t)rows * k;
        double* c0 = S + (size_t)rows * j0; double* c1 = S + (size_t)rows * (two ? j1 : j0);
        double w0 = 0.0, w1 = 0.0;
        for (int i = k + 1 + lane; i < rows; i += 32) { const double vi = v[i]; w0 += vi * c0[i]; w1 += vi * c1[i]; }
        for (int o = 16; o > 0; o >>= 1) { w0 += __shfl_xor_sync(0xffffffffu, w0, o); w1 += __shfl_xor_sync(0xffffffffu, w1, o); }
        w0 += c0[k]; w1 += c1[k];
        const double t0 = tk * w0, t1 = tk * w1;
        double ss = 0.0;                                   // (warp 0's first column is column k+1: its norm below row k+1 feeds the next reflector)
        for (int i = k + 1 + lane; i < rows; i += 32) {
            const double vi = v[i];
            const double x = c0[i] - t0 * vi;
            c0[i] = x;
            if (i > k + 1) ss += x * x;
            if (two) c1[i] -= t1 * vi;
        }
        if (lane == 0) { c0[k] -= t0; if (two) c1[k] -= t1; }
        if (next_ss && j0 == k + 1) {
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) *next_ss = ss;
        }
    }
}
// Householder QR of the rows x n block S (column-major, leading dimension rows, rows >= n) held in shared memory, by one CTA.
// On return the upper triangle holds R, the columns below the diagonal the reflector vectors (v(k) = 1 implicit), tau[k] the
// scalars.  One warp per trailing column: dot product and update of a column stay inside its warp, so a reflector costs two
// block barriers.  sh: scratch >= 34 doubles.
__device__ __forceinline__ void cta_house_qr(double* S, int rows, int n, double* tau, double* sh) {
    for (int k = 0; k < n; ++k) {
        double ss;
        if (k == 0) {
            ss = 0.0;
            for (int i = 1 + threadIdx.x; i < rows; i += blockDim.x) { const double x = S[i]; ss += x * x; }
            ss = qr_block_sum(ss, sh);
        } else ss = sh[33];                               // left by the warp that updated column k in the previous step
        const double alpha = S[k + rows * k];
        double beta = alpha, scale = 0.0, tk = 0.0;
        if (ss != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + ss), alpha);
            tk = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncthreads();                                  // everybody has read alpha
        for (int i = k + 1 + threadIdx.x; i < rows; i += blockDim.x) S[i + rows * k] *= scale;
        if (threadIdx.x == 0) { S[k + rows * k] = beta; tau[k] = tk; }
        __syncthreads();
        cta_apply_reflector(S, rows, n, k, tk, sh + 33);
        __syncthreads();
    }
}
// dorg2r in place: the reflectors left by cta_house_qr become the explicit rows x n factor Q
__device__ __forceinline__ void cta_house_formq(double* S, int rows, int n, const double* tau) {
    for (int k = n - 1; k >= 0; --k) {
        const double tk = tau[k];
        cta_apply_reflector(S, rows, n, k, tk);
        __syncthreads();
        for (int i = threadIdx.x; i < rows; i += blockDim.x) {
            double x = S[i + rows * k];
            if (i > k) x = -tk * x; else if (i == k) x = 1.0 - tk; else x = 0.0;
            S[i + rows * k] = x;
        }
        __syncthreads();
    }
}
// One level of the tree.  Node b factors `cnt` stacked source blocks:
//   level 1 : src = A (m x n, lda), block b = rows [b*m/G, (b+1)*m/G);                     Q block -> qout (ldq = m) at those rows
//   level>1 : src = the previous level's R factors (n x n each, contiguous), node b stacks R[4b .. 4b+3];  Q block (cnt*n x n) -> qout + b*4n*n
// rout + b*n*n receives the node's R (upper triangular, zeros below).  dynamic smem: (maxrows*n + n + 40) doubles.
__global__ void __launch_bounds__(TSQR_THREADS) k_tsqr_factor(const double* __restrict__ src, int level, int m, int n, int lda, int G, int nsrc,
                                                             double* __restrict__ rout, double* __restrict__ qout, int ldq) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    int rows, row0 = 0;
    if (level == 1) { row0 = (int)((long long)b * m / G); rows = (int)((long long)(b + 1) * m / G) - row0; }
    else { const int c0 = 4 * b; rows = min(4, nsrc - c0) * n; }
    double* S = smem; double* tau = S + (size_t)rows * n; double* sh = tau + n;
    if (level == 1) {
        for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; S[e] = src[(size_t)(row0 + i) + (size_t)lda * j]; }
    } else {
        for (int e = threadIdx.x; e < rows * n; e += blockDim.x) {
            const int j = e / rows, i = e - j * rows, blk = i / n, ii = i - blk * n;
            S[e] = src[((size_t)(4 * b + blk) * n + j) * n + ii];
        }
    }
    __syncthreads();
    cta_house_qr(S, rows, n, tau, sh);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int j = e / n, i = e - j * n; rout[(size_t)b * n * n + e] = (i <= j) ? S[i + rows * j] : 0.0; }
    __syncthreads();
    cta_house_formq(S, rows, n, tau);
    if (level == 1) {
        for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; qout[(size_t)(row0 + i) + (size_t)ldq * j] = S[e]; }
    } else {
        double* qb = qout + (size_t)b * 4 * n * n;                 // (4n x n, leading dimension 4n)
        for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; qb[i + (size_t)4 * n * j] = S[e]; }
    }
}
// Q(block b) <- Q1(block b) * Q2[slice] * Q3[slice] * ... (levels 2..L); qlev[l] = that level's Q blocks (4n x n each, ld 4n).
struct TsqrLevels { const double* q[8]; int count; };
__global__ void __launch_bounds__(TSQR_THREADS) k_tsqr_apply(double* __restrict__ q, int m, int n, int ldq, int G, TsqrLevels LV) {
    extern __shared__ double smem[];
    const int b = blockIdx.x;
    const int row0 = (int)((long long)b * m / G), rows = (int)((long long)(b + 1) * m / G) - row0;
    double* M = smem; double* T = M + n * n; double* QB = T + n * n;      // M, T: n x n (ld n); QB: rows x n
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) { const int j = e / rows, i = e - j * rows; QB[e] = q[(size_t)(row0 + i) + (size_t)ldq * j]; }
    int idx = b;
    for (int l = 0; l < LV.count; ++l) {
        const int node = idx / 4, slot = idx - 4 * node;
        const double* ql = LV.q[l] + (size_t)node * 4 * n * n + (size_t)slot * n;        // rows slot*n .. of the 4n x n block
        if (l == 0) {
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int j = e / n, i = e - j * n; M[e] = ql[i + (size_t)4 * n * j]; }
        } else {
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
                const int j = e / n, i = e - j * n;
                double t = 0.0;
                for (int x = 0; x < n; ++x) t += M[i + n * x] * ql[x + (size_t)4 * n * j];
                T[e] = t;
            }
            __syncthreads();
            for (int e = threadIdx.x; e < n * n; e += blockDim.x) M[e] = T[e];
        }
        __syncthreads();
        idx = node;
    }
    if (LV.count == 0) return;
    for (int e = threadIdx.x; e < rows * n; e += blockDim.x) {
        const int j = e / rows, i = e - j * rows;
        double t = 0.0;
        for (int x = 0; x < n; ++x) t += QB[i + rows * x] * M[x + n * j];
        q[(size_t)(row0 + i) + (size_t)ldq * j] = t;
    }
}
// Householder reconstruction of LAPACK's signs: modified LU (no pivoting) of the top n x n block of Q; sgn[k] = S_kk; r <- S r.
__global__ void __launch_bounds__(TSQR_THREADS) k_tsqr_sign(const double* __restrict__ q, int n, int ldq, double* __restrict__ r, double* __restrict__ sgn) {
    extern __shared__ double smem[];
    double* W = smem; double* S = W + n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) { const int j = e / n, i = e - j * n; W[e] = q[(size_t)i + (size_t)ldq * j]; }
    __syncthreads();
    for (int k = 0; k < n; ++k) {
        if (threadIdx.x == 0) { const double sk = (W[k + n * k] >= 0.0) ? -1.0 : 1.0; S[k] = sk; W[k + n * k] -= sk; }
        __syncthreads();
        const double piv = W[k + n * k];
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) W[i + n * k] /= piv;
        __syncthreads();
        for (int e = threadIdx.x; e < (n - k - 1) * (n - k - 1); e += blockDim.x) {
            const int jj = e / (n - k - 1), ii = e - jj * (n - k - 1), i = k + 1 + ii, j = k + 1 + jj;
            W[i + n * j] -= W[i + n * k] * W[k + n * j];
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
