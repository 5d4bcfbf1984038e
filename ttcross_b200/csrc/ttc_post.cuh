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

}  // namespace ttc
