// =============================================================================
// ttc_post.cuh — post-processing of the finished train on the device.
//
// ztt_quad (reference lib/dmrgg.f90:1418-1523): quadrature of the train against a rank-1 COMPLEX weight tensor, the step
// test_crs_chf.f90:153-168 / test_crs_pdf.f90:128-190 repeat for 32 frequencies to sample a characteristic function
// (SURVEY 8(f) rank 2).  The drivers convert the real train to a complex one with zero imaginary parts, so here the cores
// stay real and only the weights are complex.  All `nsets` weight sets are contracted in ONE launch, one CTA per set:
//   curr(i,k) = sum_j w_p(j) * core_p(i,j,k)       zgemv 'n' order (j ascending, from zero)           (:1469-1471)
//   prev      = prev * curr                          zgemm 'n','n' order; r(0) = 1, so prev is a row     (:1481-1483)
// in the single-rank order of the reference (first = l, last = m).  dynamic smem: 2 * Rmax doubles x 2 (prev, next) +
// 2 * nmax doubles (weights of the current core).
// =============================================================================
#pragma once
#include "ttc_device.cuh"

namespace ttc {

// wre / wim: [nsets][sum_p n(p)] (cores' weights concatenated in core order); out: [nsets] re | [nsets] im
__global__ void k_zquad(DevPlan P, int nsets, const double* __restrict__ wre, const double* __restrict__ wim, double* out_re, double* out_im, long long wstride) {
    extern __shared__ double smem[];
    const int set = blockIdx.x;
    if (set >= nsets) return;
    const int R = P.Rmax;
    double* pre = smem; double* pim = pre + R; double* nre = pim + R; double* nim = nre + R;
    double* wr = nim + R; double* wi = wr + P.nmax;
    if (threadIdx.x == 0) { pre[0] = 1.0; pim[0] = 0.0; }
    long long woff = 0;
    for (int p = 1; p <= P.d; ++p) {
        const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
        const double* a = P.arg + P.coreOff[p];
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) { wr[j] = wre[set * wstride + woff + j]; wi[j] = wim[set * wstride + woff + j]; }
        __syncthreads();
        // next(k) = sum_i prev(i) * curr(i,k), curr(i,k) = sum_j w(j) a(i,j,k): one warp per column k, lanes over rows i
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int k = wid; k < r1; k += nw) {
            double acc_re = 0.0, acc_im = 0.0;                  // C(1,k), accumulated over l = i ascending (lane 0 folds in order)
            for (int i0 = 0; i0 < r0; i0 += 32) {
                const int i = i0 + lane;
                double cre = 0.0, cim = 0.0;
                if (i < r0) {
                    const double* col = a + i + (i64)R * n * k;
#pragma unroll 4
                    for (int j = 0; j < n; ++j) { const double v = col[(i64)R * j]; cre = cre + wr[j] * v; cim = cim + wi[j] * v; }
                }
                // temp = curr(l,k); C(1,k) += temp * prev(l), l ascending: complex product (a+bi)(c+di) = (ac - bd) + (ad + bc)i
                const int cnt = min(32, r0 - i0);
                for (int u = 0; u < cnt; ++u) {
                    const double tre = __shfl_sync(FULLMASK, cre, u), tim = __shfl_sync(FULLMASK, cim, u);
                    const double are = pre[i0 + u], aim = pim[i0 + u];
                    acc_re = acc_re + (tre * are - tim * aim);
                    acc_im = acc_im + (tre * aim + tim * are);
                }
            }
            if (lane == 0) { nre[k] = acc_re; nim[k] = acc_im; }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < r1; k += blockDim.x) { pre[k] = nre[k]; pim[k] = nim[k]; }
        woff += n;
    }
    __syncthreads();
    if (threadIdx.x == 0) { out_re[set] = pre[0]; out_im[set] = pim[0]; }
}

// ----------------------------------------------------------------------------
// dtt_ijk (lib/tt.f90:630-652): one element of the train, x = core_d(:, i_d, 1); for p = d-1 .. 1: x = core_p(:, i_p, :) x.
// One warp per multi-index, the running vector in shared memory (two buffers of Rmax doubles per warp); rows of the slice
// across lanes (coalesced), the sum over the column index sequential from zero (the order of an inlined matmul).
// ----------------------------------------------------------------------------
__device__ __forceinline__ double tt_value_warp(const DevPlan& P, const int* ind, double* xa, double* xb) {
    const int lane = threadIdx.x & 31;
    const int d = P.d, R = P.Rmax;
    {
        const int r0 = P.rk[d - 1], n = P.n[d];
        const double* a = P.arg + P.coreOff[d] + (i64)R * (ind[d - 1] - 1);
        for (int i = lane; i < r0; i += 32) xa[i] = a[i];
        (void)n;
    }
    __syncwarp();
    for (int p = d - 1; p >= 1; --p) {
        const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
        const double* a = P.arg + P.coreOff[p] + (i64)R * (ind[p - 1] - 1);      // y(i,k) = a[i + R*n*k]
        for (int i = lane; i < r0; i += 32) {
            double z = 0.0;
#pragma unroll 4
            for (int k = 0; k < r1; ++k) z = z + a[i + (i64)R * n * k] * xa[k];
            xb[i] = z;
        }
        __syncwarp();
        double* t = xa; xa = xb; xb = t;
    }
    return xa[0];
}
// values[x] = train(ind[x][0..d-1]) for `count` multi-indices (1-based); dynamic smem: 2*Rmax doubles per warp
__global__ void k_tt_values(DevPlan P, long long count, const int* __restrict__ ind, double* __restrict__ values) {
    extern __shared__ double smem[];
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    double* xa = smem + (size_t)wid * 2 * P.Rmax; double* xb = xa + P.Rmax;
    for (long long x = (long long)blockIdx.x * nw + wid; x < count; x += (long long)gridDim.x * nw) {
        const double v = tt_value_warp(P, ind + x * P.d, xa, xb);
        if (lane == 0) values[x] = v;
        __syncwarp();
    }
}
// dtt_accchk (lib/dmrgg.f90:1081-1166): nlot random multi-indices ind(p) = int(u*n(p)) + 1 (irnd, rnd.f90:84-90) from the
// built-in uniform stream; a = integrand, b = train; per-CTA partials of  max|a-b| (+ sample number),  sum (a-b)^2,
// max a,  sum a^2  in `part` [gridDim.x][4] + sample numbers [gridDim.x]; the host folds them in CTA order.
template <int KIND>
__global__ void k_accchk(DevPlan P, long long nlot, unsigned long long seed, double* part, long long* argpart) {
    extern __shared__ double smem[];
    __shared__ double s_e[32], s_f[32], s_a[32], s_g[32];
    __shared__ long long s_x[32];
    const double* A = stage_aux<KIND>(P, smem);
    const int wid = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    double* xa = smem + P.auxsm + (size_t)wid * (2 * P.Rmax + P.d); double* xb = xa + P.Rmax;
    int* ind = (int*)(xb + P.Rmax);
    double einf = -1.0, efro = 0.0, ainf = 0.0, afro = 0.0; long long ex = 0;
    for (long long x = (long long)blockIdx.x * nw + wid; x < nlot; x += (long long)gridDim.x * nw) {
        for (int p = lane; p < P.d; p += 32) {
            const double u = stream_uniform(seed, 0x7fffffff, (unsigned long long)x * P.d + p);
            int v = (int)(u * P.n[p + 1]) + 1;
            if (v > P.n[p + 1]) v = P.n[p + 1];
            ind[p] = v;
        }
        __syncwarp();
        const double b = tt_value_warp(P, ind, xa, xb);
        if (lane == 0) {
            struct Idx { const int* v; __device__ __forceinline__ int operator()(int pos) const { return v[pos - 1]; } } src{ind};
            const double a = eval_src<KIND>(P, src, A);
            const double e = fabs(a - b);
            if (einf < e) { einf = e; ex = x; }
            efro = efro + (a - b) * (a - b);
            ainf = fmax(ainf, a);                          // dmax1(ainf, aval): not the absolute value, as in the reference
            afro = afro + a * a;
        }
        __syncwarp();
    }
    if (lane == 0) { s_e[wid] = einf; s_f[wid] = efro; s_a[wid] = ainf; s_g[wid] = afro; s_x[wid] = ex; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) {
            if (s_e[0] < s_e[w] || (s_e[0] == s_e[w] && s_x[w] < s_x[0])) { s_e[0] = s_e[w]; s_x[0] = s_x[w]; }
            s_f[0] += s_f[w]; s_a[0] = fmax(s_a[0], s_a[w]); s_g[0] += s_g[w];
        }
        part[4 * blockIdx.x] = s_e[0]; part[4 * blockIdx.x + 1] = s_f[0]; part[4 * blockIdx.x + 2] = s_a[0]; part[4 * blockIdx.x + 3] = s_g[0];
        argpart[blockIdx.x] = s_x[0];
    }
}

// ----------------------------------------------------------------------------
// dtt_svd (lib/tt.f90:307-368): TT rounding.  After dtt_ort the unfolding M = core_k viewed as r(k-1) x (n(k) r(k)) is
// short and fat; its SVD is taken as  M^T = Q R  (tall-skinny QR, k_qr_panel)  and  R^T = U S W^T  (one-sided Jacobi on the
// small r x r matrix inside one CTA, below), so  M = U S (Q W)^T.  The truncation rule is mat.f90's `chop` (host, the r
// singular values), U S goes into core k-1 and the kept rows of (Q W)^T become core k.
// ----------------------------------------------------------------------------
// T(jk + nn*i) = core(i + ld*jk): the transposed unfolding, contiguous
__global__ void k_svd_pack_t(const double* core, int mm, long long nn, int ld, double* T) {
    const long long tot = (long long)mm * nn;
    for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (long long)gridDim.x * blockDim.x) {
        const long long jk = x / mm; const int i = (int)(x - jk * mm);
        T[jk + nn * i] = core[i + (long long)ld * jk];
    }
}
// One-sided Jacobi SVD of A = R^T (m x m, R upper triangular from the QR, leading dimension m): A J = B with orthogonal
// columns, B = U S, W = J.  One CTA; a warp per column pair, pairs of a round from the round-robin (tournament) schedule.
// Output: U (m x m), s (m, descending), W (m x m), all column-major with leading dimension m.
// dynamic smem: 2*m*m doubles (A, W) + 2*m doubles (norms, scratch) + 2*m ints (perm, order)
__global__ void k_svd_small(const double* __restrict__ R, int m, double* U, double* sv, double* Wout) {
    extern __shared__ double smem[];
    double* A = smem; double* W = A + (size_t)m * m; double* nrm = W + (size_t)m * m; double* tmp = nrm + m;
    int* perm = (int*)(tmp + m); int* ord = perm + m + 1;
    __shared__ int s_rot;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) { const int j = e / m, i = e - j * m; A[e] = R[j + (size_t)m * i]; W[e] = (i == j) ? 1.0 : 0.0; }   // A = R^T
    const int me = m + (m & 1);                       // even number of players (a dummy sits out)
    for (int e = threadIdx.x; e < me; e += blockDim.x) perm[e] = e;
    __syncthreads();
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (threadIdx.x == 0) s_rot = 0;
        __syncthreads();
        for (int round = 0; round < me - 1; ++round) {
            for (int pr = wid; pr < me / 2; pr += nw) {
                int p = perm[pr], q = perm[me - 1 - pr];
                if (p >= m || q >= m) continue;
                if (p > q) { const int t = p; p = q; q = t; }
                double a = 0.0, b = 0.0, g = 0.0;
                for (int i = lane; i < m; i += 32) { const double x = A[i + m * p], y = A[i + m * q]; a += x * x; b += y * y; g += x * y; }
                for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); g += __shfl_xor_sync(0xffffffffu, g, o); }
                if (fabs(g) > 1e-15 * sqrt(a * b) && g != 0.0) {
                    const double zeta = (b - a) / (2.0 * g);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                    for (int i = lane; i < m; i += 32) {
                        const double x = A[i + m * p], y = A[i + m * q];
                        A[i + m * p] = c * x - sn * y; A[i + m * q] = sn * x + c * y;
                        const double wx = W[i + m * p], wy = W[i + m * q];
                        W[i + m * p] = c * wx - sn * wy; W[i + m * q] = sn * wx + c * wy;
                    }
                    if (lane == 0) s_rot = 1;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) { const int last = perm[me - 1]; for (int e = me - 1; e > 1; --e) perm[e] = perm[e - 1]; perm[1] = last; }
            __syncthreads();
        }
        if (!s_rot) break;
        __syncthreads();
    }
    // singular values = column norms, descending order, U = normalised columns
    for (int j = wid; j < m; j += nw) {
        double a = 0.0;
        for (int i = lane; i < m; i += 32) { const double x = A[i + m * j]; a += x * x; }
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) nrm[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < m; j += blockDim.x) {        // rank of column j in the descending order (ties by index)
        int rnk = 0;
        for (int u = 0; u < m; ++u) rnk += (nrm[u] > nrm[j]) || (nrm[u] == nrm[j] && u < j);
        ord[rnk] = j;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
        const int jo = e / m, i = e - jo * m, j = ord[jo];
        const double sj = nrm[j];
        U[e] = (sj > 0.0) ? A[i + m * j] / sj : ((i == jo) ? 1.0 : 0.0);
        Wout[e] = W[i + m * j];
    }
    for (int jo = threadIdx.x; jo < m; jo += blockDim.x) sv[jo] = nrm[ord[jo]];
}
// tmp(x + kk*j) = sum_l prevcore(x, l) * U(l, j) * s(j), x < kk, j < rr  (dgemm 'n','n' of tt.f90:342 with U scaled first, :338-340)
__global__ void k_svd_apply_left(const double* prev, long long kk0, int nprev, int ldp, const double* U, const double* sn, int mm, int rr, double* tmp) {
    // prev is the padded core k-1: element (i, j, l) at i + ldp*(j + nprev*l); x = i + r(k-2)*j enumerates its rows
    const long long tot = kk0 * rr;
    const int r2 = (int)(kk0 / nprev);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e / kk0); const long long x = e - (long long)j * kk0;
        const int jj = (int)(x / r2), i = (int)(x - (long long)jj * r2);
        double t = 0.0;
        for (int l = 0; l < mm; ++l) t = t + (U[l + (size_t)mm * j] * sn[j]) * prev[i + (long long)ldp * (jj + (long long)nprev * l)];
        tmp[e] = t;
    }
}
__global__ void k_svd_store_left(const double* tmp, double* prev, long long kk0, int nprev, int ldp, int rr) {
    const long long tot = kk0 * rr;
    const int r2 = (int)(kk0 / nprev);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(e / kk0); const long long x = e - (long long)j * kk0;
        const int jj = (int)(x / r2), i = (int)(x - (long long)jj * r2);
        prev[i + (long long)ldp * (jj + (long long)nprev * j)] = tmp[e];
    }
}
// core_k(i + ld*jk) = V(i, jk) = sum_j W(j, i) * Q(jk, j), i < rr  (the kept rows of (Q W)^T)
__global__ void k_svd_apply_right(const double* Q, long long nn, const double* W, int mm, int rr, double* core, int ld) {
    const long long tot = nn * rr;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (long long)gridDim.x * blockDim.x) {
        const long long jk = e / rr; const int i = (int)(e - jk * rr);
        double t = 0.0;
        for (int j = 0; j < mm; ++j) t = t + W[j + (size_t)mm * i] * Q[jk + nn * j];
        core[i + (long long)ld * jk] = t;
    }
}

}  // namespace ttc
