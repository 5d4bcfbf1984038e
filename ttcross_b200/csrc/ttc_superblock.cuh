// =============================================================================
// ttc_superblock.cuh — the full-pivoting superblock of dtt_dmrgg (pivoting = -1, dmrgg.f90:341-396), tiled.
//
//   a(i,j,k,q) = f(left pivot i, j, k, right pivot q)                      r(p-1)*n(p) rows  x  n(p+1)*r(p+1) columns
//   b          = a - col(p)(:,:,1:r) * row(p+1)(1:r,:,:)                   dgemm 'n','n', K = r(p)   (dmrgg.f90:383-386)
//   two first-index argmaxes (idamax, dmrgg.f90:373-396)
//
// This is the one roofline-sized kernel of the path: 67.6 M evaluations and a 8224 x 8224 x 32 contraction
// per bond visit at the C_10 shape.  Layout of the work:
//   * a CTA owns TM = 256 consecutive rows (one per thread) and walks a range of columns in tiles of TN = 32;
//   * the K x TM slab of the column factor stays in shared memory for the CTA's whole life; the K x TN slab of the row
//     factor and the per-column integrand state are rebuilt per tile by a few threads;
//   * every thread carries 8 columns at a time in registers: per step of the contraction one conflict-free shared load
//     (its row of the column factor) and four 16-byte broadcast loads feed 16 FP64 instructions, which is what lets the
//     FP64 pipe (64 lanes/SM), not the shared-memory port, set the pace.
//   * Ising C (test_crs_ising.f90:196-217) is evaluated from per-row and per-column partial recurrences: the prefix
//     recurrence (w) over the left positions and the suffix recurrence (v) over the right positions are advanced ONCE per
//     row / per column and continued per element — the same operations in the same order as the reference loop, so the
//     result is bit-identical, at 3m+3 instead of 5m+3 flops per element.
// FMA = 0 keeps the reference arithmetic (separate multiply and add): results are bit-identical to the CPU
// oracle.  FMA = 1 contracts the residual update into DFMA (measurement of the FP64 ceiling; not used by ttc_dmrgg).
// =============================================================================
#pragma once
#include "ttc_device.cuh"

namespace ttc {

#ifndef TTC_SB_TM
#define TTC_SB_TM 256
#endif
#ifndef TTC_SB_CN
#define TTC_SB_CN 8
#endif
#ifndef TTC_SB_TN
#define TTC_SB_TN 32
#endif
#ifndef TTC_SB_MINB
#define TTC_SB_MINB 2
#endif
constexpr int SB_TM = TTC_SB_TM;     // rows per CTA = threads per CTA
constexpr int SB_CN = TTC_SB_CN;     // columns a thread carries in registers
constexpr int SB_TN = TTC_SB_TN;     // columns per tile
constexpr int SB_MINB = TTC_SB_MINB; // CTAs per SM the register allocation aims at
constexpr int SB_MAXL = 8;       // left positions (+ the free mode) kept in registers on the fast Ising-C path

__host__ __device__ __forceinline__ size_t sb_tile_doubles(int Rmax, int d) {
    // Ct[Rmax*TM] | Rt[Rmax*TN] | XK WK VK VV [4*TN] | XRt WRt [2*d*TN]
    return (size_t)Rmax * SB_TM + (size_t)Rmax * SB_TN + 4 * SB_TN + 2 * (size_t)d * SB_TN + 2;   // + alignment slack
}
// FMA = 2 (fast mode): the residual b = a - col * row goes through the FP64 tensor-core path (mma.sync.m8n8k4.f64, SASS DMMA).
// The evaluated tile a (TM x TN) is parked in shared memory (Ft, column-major, padded), then every warp owns 32 rows x 32
// columns = 4 x 4 fragments: accumulators start from a, 8 k-steps of 4 contract K = 32.  Not bit-identical to the reference
// (FMA contraction, k summed in the tensor core's order): pivot agreement with parity mode is reported by the tests / bench.
constexpr int SB_TMP = SB_TM + 2;                // padded leading dimension of Ft
__host__ __device__ __forceinline__ size_t sb_mma_doubles() { return (size_t)SB_TN * SB_TMP; }
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int KIND, int STORE, int FMA>
__global__ void __launch_bounds__(SB_TM, FMA == 2 ? 1 : SB_MINB) k_superblock_t(DevPlan P, int dir, int pp, int fixed_bond, int fixed_v, double* a_out, Partial* probe_out) {
    tl_stamp(P, 35);
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    __shared__ int s_last;
    const int v = (fixed_bond > 0) ? fixed_v : P.v0 + blockIdx.z;
    Dims D;
    if (fixed_bond > 0) {
        D.active = 1; D.p = fixed_bond; D.r0 = P.rk[D.p - 1]; D.r1 = P.rk[D.p]; D.r2 = P.rk[D.p + 1];
        D.n1 = P.n[D.p]; D.n2 = P.n[D.p + 1];
    } else {
        if (P.ctrl->ready) return;
        D = load_dims(P, v, dir, pp);
    }
    if (!D.active) return;
    const int nrb = (D.r0 * D.n1 + SB_TM - 1) / SB_TM;      // row blocks that exist at the current ranks (the grid is sized for the capacity)
    if ((int)blockIdx.x >= nrb) return;
    const int K = D.r1;
    const double* A = stage_aux<KIND>(P, smem);
    double* stg = smem + P.auxsm;
    const Stage S = stage_bond(P, stg, D.p - 1, D.r0, D.p, D.p + 1, D.p + 1, D.r2);
    double* Ct = stg + P.stage_max;                    // Ct[s*TM + row]
    if ((size_t)(Ct - smem) & 1) ++Ct;                 // 16-byte alignment for the double2 loads of the row-factor tile
    double* Rt = Ct + (size_t)P.Rmax * SB_TM;          // Rt[s*TN + col]
    double* XK = Rt + (size_t)P.Rmax * SB_TN; double* WK = XK + SB_TN; double* VK = WK + SB_TN; double* VV = VK + SB_TN;
    double* XRt = VV + SB_TN;                          // XRt[t*TN + col]
    double* WRt = XRt + (size_t)P.d * SB_TN;
    double* Ft = WRt + (size_t)P.d * SB_TN;            // (FMA = 2 only) evaluated tile, Ft[col * TMP + row]
    const int K4 = (D.r1 + 3) & ~3;                    // K rounded up to the k-step of the DMMA (extra rows are zero)
    const double* colp = P.col + P.coreOff[D.p];
    const double* rowp = P.rowT + P.coreOff[D.p + 1];
    const i64 cs = (i64)P.Rmax * D.n1;
    const i64 rs = (i64)D.n2 * P.Rmax;
    const int m1 = D.r0 * D.n1, ncols = D.n2 * D.r2;
    const int nl = D.p - 1, nr = P.d - D.p - 1;        // left / right positions around the two free modes
    const bool fastc = (KIND == KIND_ISING) && P.ising_id == 1 && nl <= SB_MAXL;

    // ---- this thread's row
    const int row = blockIdx.x * SB_TM + threadIdx.x;
    const bool rvalid = row < m1;
    const int rowc = rvalid ? row : m1 - 1;
    const int ri = rowc % D.r0 + 1, rj = rowc / D.r0 + 1;
    {
        const double* cr = colp + (ri - 1) + (i64)P.Rmax * (rj - 1);
        for (int s0 = 0; s0 < K; s0 += 8) {
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = cr[(i64)min(s0 + u, K - 1) * cs];
#pragma unroll
            for (int u = 0; u < 8; ++u) if (s0 + u < K) Ct[(s0 + u) * SB_TM + threadIdx.x] = t[u];
        }
        if (FMA == 2) for (int s1 = K; s1 < K4; ++s1) Ct[s1 * SB_TM + threadIdx.x] = 0.0;
    }
    double xl[SB_MAXL], wl[SB_MAXL];
    double xj = 0.0, wj = 0.0, wk_row = 1.0, w_row = 1.0;
    if (fastc) {
        xj = S.NX[rj - 1]; wj = S.NW[rj - 1];
#pragma unroll
        for (int pos = 0; pos < SB_MAXL; ++pos) {
            xl[pos] = (pos < nl) ? S.XL[pos * D.r0 + (ri - 1)] : 1.0;
            wl[pos] = (pos < nl) ? S.WL[pos * D.r0 + (ri - 1)] : 1.0;
        }
#pragma unroll
        for (int pos = 0; pos < SB_MAXL; ++pos) if (pos < nl) { wk_row = wk_row * xl[pos]; w_row = w_row + wk_row; }
        wk_row = wk_row * xj; w_row = w_row + wk_row;
    }
    // ---- this CTA's column tiles
    const int ntiles = (ncols + SB_TN - 1) / SB_TN;
    const int tpc = (ntiles + gridDim.y - 1) / gridDim.y;
    const int t_begin = blockIdx.y * tpc, t_end = min(ntiles, t_begin + tpc);
    Partial braw = amax_init(), bres = amax_init();
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int c0 = tile * SB_TN;
        __syncthreads();                                  // previous tile fully consumed
        for (int e = threadIdx.x; e < K * SB_TN; e += blockDim.x) {
            const int s = e / SB_TN, c = e - s * SB_TN;
            const int kq = min(c0 + c, ncols - 1);
            Rt[e] = rowp[kq + (i64)s * rs];               // (k-1) + n2*(q-1) = kq
        }
        if (FMA == 2) for (int e = K * SB_TN + threadIdx.x; e < K4 * SB_TN; e += blockDim.x) Rt[e] = 0.0;
        if (fastc && threadIdx.x < SB_TN) {
            const int c = threadIdx.x;
            const int kq = min(c0 + c, ncols - 1);
            const int q = kq / D.n2 + 1, k = kq % D.n2 + 1;
            double vk = 1.0, vv = 1.0;
            for (int t = nr - 1; t >= 0; --t) {
                const double xr = S.XR[t * D.r2 + (q - 1)];
                XRt[t * SB_TN + c] = xr; WRt[t * SB_TN + c] = S.WR[t * D.r2 + (q - 1)];
                vk = vk * xr; vv = vv + vk;
            }
            const double xk = S.NX2[k - 1];
            vk = vk * xk; vv = vv + vk;
            XK[c] = xk; WK[c] = S.NW2[k - 1]; VK[c] = vk; VV[c] = vv;
        }
        __syncthreads();
#pragma unroll 1
        for (int cb = 0; cb < SB_TN; cb += SB_CN) {
            double f[SB_CN];
            if (fastc) {
                double vk[SB_CN], vv[SB_CN], wk[SB_CN], w[SB_CN];
#pragma unroll
                for (int c = 0; c < SB_CN; ++c) {
                    vk[c] = VK[cb + c] * xj; vv[c] = VV[cb + c] + vk[c];
                    wk[c] = wk_row * XK[cb + c]; w[c] = w_row + wk[c];
                }
#pragma unroll
                for (int pos = SB_MAXL - 1; pos >= 0; --pos) if (pos < nl) {
#pragma unroll
                    for (int c = 0; c < SB_CN; ++c) { vk[c] = vk[c] * xl[pos]; vv[c] = vv[c] + vk[c]; }
                }
                for (int t = 0; t < nr; ++t) {
#pragma unroll
                    for (int c = 0; c < SB_CN; ++c) { wk[c] = wk[c] * XRt[t * SB_TN + cb + c]; w[c] = w[c] + wk[c]; }
                }
#pragma unroll
                for (int c = 0; c < SB_CN; ++c) { const double b = 1.0 / (vv[c] * w[c]); f[c] = 2 * b; }
#pragma unroll
                for (int pos = 0; pos < SB_MAXL; ++pos) if (pos < nl) {
#pragma unroll
                    for (int c = 0; c < SB_CN; ++c) f[c] = f[c] * wl[pos];
                }
#pragma unroll
                for (int c = 0; c < SB_CN; ++c) { f[c] = f[c] * wj; f[c] = f[c] * WK[cb + c]; }
                for (int t = 0; t < nr; ++t) {
#pragma unroll
                    for (int c = 0; c < SB_CN; ++c) f[c] = f[c] * WRt[t * SB_TN + cb + c];
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < SB_CN; ++c) {
                    const int kq = min(c0 + cb + c, ncols - 1);
                    StagedVals sv = S.point(ri, rj, kq % D.n2 + 1, kq / D.n2 + 1);
                    f[c] = eval_point<KIND>(P, sv, A);
                }
            }
            if (FMA == 2) {
                // fast mode: park the evaluated values; the residual of the whole tile is taken by the warps below
#pragma unroll
                for (int c = 0; c < SB_CN; ++c) {
                    Ft[(cb + c) * SB_TMP + threadIdx.x] = f[c];
                    const int kq = c0 + cb + c;
                    if (rvalid && kq < ncols) {
                        const i64 x = (i64)row + (i64)m1 * kq;
                        if (STORE) a_out[x] = f[c];
                        amax_take(braw, f[c], x);
                    }
                }
                continue;
            }
            double res[SB_CN];
#pragma unroll
            for (int c = 0; c < SB_CN; ++c) res[c] = f[c];
#pragma unroll 8
            for (int s = 0; s < K; ++s) {
                const double cv = Ct[s * SB_TM + threadIdx.x];
                const double2* rp = reinterpret_cast<const double2*>(Rt + s * SB_TN + cb);
#pragma unroll
                for (int c2 = 0; c2 < SB_CN / 2; ++c2) {
                    const double2 rv = rp[c2];
                    if (FMA == 1) {
                        res[2 * c2] = fma(-rv.x, cv, res[2 * c2]);
                        res[2 * c2 + 1] = fma(-rv.y, cv, res[2 * c2 + 1]);
                    } else {
                        res[2 * c2] = res[2 * c2] + (-rv.x) * cv;
                        res[2 * c2 + 1] = res[2 * c2 + 1] + (-rv.y) * cv;
                    }
                }
            }
            if (rvalid) {
#pragma unroll
                for (int c = 0; c < SB_CN; ++c) {
                    const int kq = c0 + cb + c;
                    if (kq < ncols) {
                        const i64 x = (i64)row + (i64)m1 * kq;
                        if (STORE) a_out[x] = f[c];
                        amax_take(braw, f[c], x);        // this thread meets its elements in ascending linear index:
                        amax_take(bres, res[c], x);      // strict '>' keeps the first maximum (idamax)
                    }
                }
            }
        }
        if (FMA == 2) {
            __syncthreads();                               // the tile is complete in Ft
            const int lane = threadIdx.x & 31, wrow = (threadIdx.x >> 5) * 32, g = lane >> 2, tq = lane & 3;
            double acc[4][4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    acc[mi][ni][0] = Ft[(ni * 8 + 2 * tq) * SB_TMP + wrow + mi * 8 + g];
                    acc[mi][ni][1] = Ft[(ni * 8 + 2 * tq + 1) * SB_TMP + wrow + mi * 8 + g];
                }
            for (int kk = 0; kk < K4; kk += 4) {
                double af[4], bf[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) af[mi] = -Ct[(kk + tq) * SB_TM + wrow + mi * 8 + g];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) bf[ni] = Rt[(kk + tq) * SB_TN + ni * 8 + g];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(acc[mi][ni], af[mi], bf[ni]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int e = 0; e < 2; ++e)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) {
                        const int r_ = blockIdx.x * SB_TM + wrow + mi * 8 + g, kq = c0 + ni * 8 + 2 * tq + e;
                        if (r_ < m1 && kq < ncols) {
                            Partial cand; cand.absv = fabs(acc[mi][ni][e]); cand.val = acc[mi][ni][e]; cand.idx = (i64)r_ + (i64)m1 * kq;
                            amax_merge(bres, cand);      // (fragment order is not index order: ties go to the smaller index explicitly)
                        }
                    }
        }
    }
    // ---- grid-wide first-index argmax: last CTA (of this virtual rank) folds the partials
    __threadfence();
    braw = amax_block(braw, shp);
    bres = amax_block(bres, shp);
    const int nblk = nrb * gridDim.y, bid = blockIdx.x + nrb * blockIdx.y;
    Partial* part = P.part + (i64)v * 2 * GMAX;
    if (threadIdx.x == 0) {
        part[bid] = braw;
        part[GMAX + bid] = bres;
        __threadfence();
        const unsigned t = atomicAdd(P.tickets + v, 1u);
        s_last = (t == (unsigned)nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const volatile Partial* vp = part;
        Partial a = amax_init(), b2 = amax_init();
        for (int x = threadIdx.x; x < nblk; x += blockDim.x) {
            Partial t1, t2;
            t1.absv = vp[x].absv; t1.val = vp[x].val; t1.idx = vp[x].idx;
            t2.absv = vp[GMAX + x].absv; t2.val = vp[GMAX + x].val; t2.idx = vp[GMAX + x].idx;
            amax_merge(a, t1); amax_merge(b2, t2);
        }
        braw = amax_block(a, shp);
        bres = amax_block(b2, shp);
    }
    if (threadIdx.x == 0) {
        P.tickets[v] = 0;
        if (probe_out) { probe_out[0] = braw; probe_out[1] = bres; return; }
        VState& St = P.st[v];
        St.amax = fmax(St.amax, braw.absv);
        const i64 x = bres.idx;
        const i64 kq = x / m1; const int ij = (int)(x - kq * m1);
        St.qq = (int)(kq / D.n2) + 1; St.kk = (int)(kq % D.n2) + 1;
        St.jj = ij / D.r0 + 1; St.ii = ij % D.r0 + 1;
        St.pivot = bres.val;
        St.done = 1; St.havecol = 1; St.haverow = 1; St.crs = 0; St.upd = 0;
        St.neval += (i64)m1 * ncols;
    }
}

// FP64 pipe ceiling of this device, measured: 8 independent chains per thread, 16 warps per CTA, 2 CTAs per SM.
// fma = 1: DFMA (2 flops per instruction); fma = 0: separate DMUL + DADD (the reference arithmetic).
template <int FMA>
__global__ void __launch_bounds__(512) k_fp64_peak(double* out, int iters, double seed) {
    double a[8], b = seed, c = 1.0 - seed * 1e-9;
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = seed + u + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (FMA) a[u] = fma(a[u], c, b);
                else { a[u] = a[u] * c; a[u] = a[u] + b; }
            }
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += a[u];
    if (sum == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = sum;   // keeps the chains alive
}

}  // namespace ttc
