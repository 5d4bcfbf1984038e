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
