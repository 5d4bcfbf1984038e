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

}  // namespace ttc
