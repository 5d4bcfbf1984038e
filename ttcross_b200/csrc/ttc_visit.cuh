// =============================================================================
// ttc_visit.cuh — one thread-block CLUSTER per virtual rank runs all bond visits of a sweep.
//
// A bond visit of the rook search (dmrgg.f90:410-758) is a chain of 6-7 tiny dependent steps
// (lottery -> up to 2*piv cross fibers -> accept -> rank-1 update), each a few thousand integrand
// evaluations followed by a first-index argmax.  As separate kernels every step pays a launch, a
// grid ramp, a chain of dependent global loads for the geometry, and a last-CTA reduction through
// HBM (measured on B200: ~10 us per fiber step, ~100 us per visit).  Here the whole visit list of a
// virtual rank lives in ONE kernel: the cluster's CTAs split every fiber, argmax partials are
// exchanged through distributed shared memory, and the steps are separated by the hardware
// cluster barrier.  Scalar state (pivot candidate, amax, neval, RNG position) is replicated in
// every CTA's shared memory and advanced identically from the same reduced partials.
//
// Arithmetic, summation orders and tie-breaks are those of ttc_device.cuh (bit-identical results).
// Data that other CTAs of the cluster may have written earlier in the same kernel (ranks, index
// tables, packed LUs, factor cores, fiber buffers) is read only after a barrier.cluster, whose
// release/acquire semantics (and L1 invalidation) make ordinary loads see it (see LDF below).
// =============================================================================
#pragma once
#include "ttc_device.cuh"
#include <cooperative_groups.h>

namespace ttc {
namespace cg = cooperative_groups;

constexpr int MAXCS = 16;     // largest cluster the hardware allows (non-portable size)

struct VisitShared {
    Partial xres[2][MAXCS];       // [fold parity][CTA rank]: every CTA receives every CTA's residual partial (pushed)
    double xraw[2][MAXCS];        // same for the largest |f|
    Partial red;
    double redraw;
    Partial shp[32];
    double shr[32];
    int nz[2];
    VState S;
    int r0, r1, r2;
};

// Loads of data that other CTAs of the cluster may have written earlier in this kernel (ranks, index tables, packed LUs,
// factor cores, fiber buffers).  Every such write is separated from its next read by barrier.cluster (cluster.sync):
// arrive.release / wait.acquire at cluster scope orders them, and the hardware invalidates the SM's L1 at the barrier
// (CCTL.IVALL), so ordinary cached loads are correct; -DTTC_STRONG_FACTOR_LOADS turns them into ld.global.cg.
#ifdef TTC_STRONG_FACTOR_LOADS
#define LDF(p) __ldcg(p)
#else
#define LDF(p) (*(p))
#endif
// ---- L2 (cache-global) variants of the residual / staging helpers of ttc_device.cuh
__device__ __forceinline__ double resid_axpy_cg(double f, const double* base, i64 stride, const double* xs, int r) {
    double res = f;
    int s0 = 0;
    for (; s0 + RU <= r; s0 += RU) {
        double a[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) a[u] = LDF(base + (s0 + u) * stride);
#pragma unroll
        for (int u = 0; u < RU; ++u) res = res + (-xs[s0 + u]) * a[u];
    }
    for (; s0 < r; ++s0) res = res + (-xs[s0]) * LDF(base + s0 * stride);
    return res;
}
__device__ __forceinline__ double resid_dot_cg(double f, const double* base, i64 stride, const double* xs, int r) {
    double t = 0.0;
    int s0 = 0;
    for (; s0 + RU <= r; s0 += RU) {
        double a[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) a[u] = LDF(base + (s0 + u) * stride);
#pragma unroll
        for (int u = 0; u < RU; ++u) t = t + a[u] * xs[s0 + u];
    }
    for (; s0 < r; ++s0) t = t + LDF(base + s0 * stride) * xs[s0];
    return f + (-t);
}
__device__ __forceinline__ double resid_ddot2_cg(double f, const double* c, i64 cs, const double* r, i64 rs, int r1) {
    double t = 0.0;
    int s0 = 0;
    for (; s0 + RU <= r1; s0 += RU) {
        double a[RU], b[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) { a[u] = LDF(c + (s0 + u) * cs); b[u] = LDF(r + (s0 + u) * rs); }
#pragma unroll
        for (int u = 0; u < RU; ++u) t = t + a[u] * b[u];
    }
    for (; s0 < r1; ++s0) t = t + LDF(c + s0 * cs) * LDF(r + s0 * rs);
    return f - t;
}
// residuals with the first RP factor values loaded BEFORE the evaluation (their L2 round trip hides behind it)
constexpr int RP = 16;
struct Pref { double a[RP]; };
__device__ __forceinline__ void pref_load(Pref& pf, const double* base, i64 stride, int r) {
#pragma unroll
    for (int u = 0; u < RP; ++u) pf.a[u] = LDF(base + min(u, r - 1) * stride);   // clamped, never predicated: all in flight at once
}
__device__ __forceinline__ double resid_axpy_pf(double f, const Pref& pf, const double* base, i64 stride, const double* xs, int r) {
    double res = f;
#pragma unroll
    for (int u = 0; u < RP; ++u) if (u < r) res = res + (-xs[u]) * pf.a[u];
    for (int s0 = RP; s0 < r; s0 += RP) {
        double a[RP];
#pragma unroll
        for (int u = 0; u < RP; ++u) a[u] = LDF(base + min(s0 + u, r - 1) * stride);
#pragma unroll
        for (int u = 0; u < RP; ++u) if (s0 + u < r) res = res + (-xs[s0 + u]) * a[u];
    }
    return res;
}
__device__ __forceinline__ double resid_dot_pf(double f, const Pref& pf, const double* base, i64 stride, const double* xs, int r) {
    double t = 0.0;
#pragma unroll
    for (int u = 0; u < RP; ++u) if (u < r) t = t + pf.a[u] * xs[u];
    for (int s0 = RP; s0 < r; s0 += RP) {
        double a[RP];
#pragma unroll
        for (int u = 0; u < RP; ++u) a[u] = LDF(base + min(s0 + u, r - 1) * stride);
#pragma unroll
        for (int u = 0; u < RP; ++u) if (s0 + u < r) t = t + a[u] * xs[s0 + u];
    }
    return f + (-t);
}
// stage_bond of ttc_device.cuh with the index tables read through L2 (they grow during the kernel)
__device__ __forceinline__ Stage stage_bond_cg(const DevPlan& P, double* sm, int pl, int rl, int c1, int c2, int pr, int rr) {
    Stage S;
    const int n1 = P.n[c1], n2 = c2 ? P.n[c2] : 0;
    const int nl = pl, nr = P.d - pr;
    const bool hasw = (P.kind == KIND_ISING);
    const int nwoff = P.n[1];
    double* NX = sm; double* NW = NX + n1; double* NX2 = NW + n1; double* NW2 = NX2 + n2;
    double* XL = NW2 + n2; double* WL = XL + nl * rl; double* XR = WL + nl * rl; double* WR = XR + nr * rr;
    for (int x = threadIdx.x; x < n1; x += blockDim.x) { NX[x] = P.par[x]; NW[x] = hasw ? P.par[nwoff + x] : 0.0; }
    for (int x = threadIdx.x; x < n2; x += blockDim.x) { NX2[x] = P.par[x]; NW2[x] = hasw ? P.par[nwoff + x] : 0.0; }
    const int* L = P.Lidx + P.offL[pl];
    for (int x = threadIdx.x; x < nl * rl; x += blockDim.x) {
        int pos = x / rl, t = x - pos * rl;
        int idx = LDF(L + (i64)pos * P.Rmax + t);
        XL[x] = P.par[idx - 1]; WL[x] = hasw ? P.par[nwoff + idx - 1] : 0.0;
    }
    const int* R = P.Ridx + P.offR[pr];
    for (int x = threadIdx.x; x < nr * rr; x += blockDim.x) {
        int pos = x / rr, t = x - pos * rr;
        int idx = LDF(R + (i64)pos * P.Rmax + t);
        XR[x] = P.par[idx - 1]; WR[x] = hasw ? P.par[nwoff + idx - 1] : 0.0;
    }
    __syncthreads();
    S.NX = NX; S.NW = NW; S.NX2 = NX2; S.NW2 = NW2; S.XL = XL; S.WL = WL; S.XR = XR; S.WR = WR;
    S.nl = nl; S.rl = rl; S.nr = nr; S.rr = rr; S.hask = c2 ? 1 : 0;
    return S;
}
// Packed LU -> transposed shared-memory table, one warp per row of the packed block with the rows of a warp in flight together
// (the block is read once per use; a dependent load per loop trip would cost one L2 round trip each).
__device__ __forceinline__ void stage_luar_cg(const double* g, int r, double* T) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll 4
    for (int s = wid; s < r; s += nw)
        for (int u = lane; u < s; u += 32) T[u * r + s] = LDF(g + (i64)s * s + u);
}
__device__ __forceinline__ void stage_lual_cg(const double* g, int r, double* T, double* dinv) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll 4
    for (int c = wid; c < r; c += nw) {
        const double* gc = g + (i64)(c + 1) * (c + 1) - (c + 1);
        for (int u = lane; u <= c; u += 32) {
            const double val = LDF(gc + u);
            if (u < c) T[u * r + c] = val; else dinv[c] = 1.0 / val;
        }
    }
}

// Fold over the whole cluster: the first-index argmax of the residuals (idamax) and the largest |f| (only its magnitude
// is used: amax).  Every thread of every CTA returns with the folded values.  Each CTA PUSHES its partial into every
// CTA's shared memory before the one cluster barrier of the call, so nobody reads remote memory after it; the exchange
// area is double-buffered by call parity (a CTA two folds ahead would have had to pass the barrier in between).
__device__ __forceinline__ void cluster_fold(cg::cluster_group& cl, VisitShared& sh, int& phase, double& rawmax, Partial& res) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    res = amax_warp(res);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rawmax = fmax(rawmax, __shfl_down_sync(FULLMASK, rawmax, o));
    if (lane == 0) { sh.shp[w] = res; sh.shr[w] = rawmax; }
    __syncthreads();
    const int buf = phase & 1;
    ++phase;
    const int cs = (int)cl.num_blocks(), me = (int)cl.block_rank();
    if (w == 0) {
        Partial a = (lane < nw) ? sh.shp[lane] : amax_init();
        double r = (lane < nw) ? sh.shr[lane] : -1.0;
        a = amax_warp(a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(FULLMASK, r, o));
        a.absv = __shfl_sync(FULLMASK, a.absv, 0); a.val = __shfl_sync(FULLMASK, a.val, 0); a.idx = __shfl_sync(FULLMASK, a.idx, 0);
        r = __shfl_sync(FULLMASK, r, 0);
        if (lane < cs) {
            VisitShared* dst = cl.map_shared_rank(&sh, lane);
            dst->xres[buf][me] = a;
            dst->xraw[buf][me] = r;
        }
    }
    cl.sync();                  // barrier.cluster arrive.release / wait.acquire: the pushes and the fiber stores are visible
    if (w == 0) {
        Partial b = (lane < cs) ? sh.xres[buf][lane] : amax_init();
        double r = (lane < cs) ? sh.xraw[buf][lane] : -1.0;
        b = amax_warp(b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r = fmax(r, __shfl_down_sync(FULLMASK, r, o));
        if (lane == 0) { sh.red = b; sh.redraw = r; }
    }
    __syncthreads();
    res = sh.red;
    rawmax = sh.redraw;
}

// ----------------------------------------------------------------------------
// Close of a sweep inside the sweep's last kernel (single process).  The exit test of dmrgg.f90:1010-1019 needs no
// quadrature value, so the last CTA / cluster to finish (arrival counter) runs k_sweep_log's body right away: one launch
// less on the critical path.  Values other CTAs produced in this kernel are read through volatile / ld.global.cg.
// ----------------------------------------------------------------------------
__device__ __forceinline__ void sweep_close_fused(const DevPlan& P, int maxrank) {
    volatile VState* st = P.st;
    const int it = P.ctrl->it;
    __shared__ unsigned long long s_ne;
    __shared__ double s_amax, s_pmax, s_pmin;
    if (threadIdx.x == 0) { s_ne = 0ULL; s_amax = st[0].amax; s_pmax = st[0].pivotmax; s_pmin = st[0].pivotmin; }
    for (int x = threadIdx.x; x <= P.d; x += blockDim.x) {
        const int r = __ldcg(P.rk + x);
        P.rklog[(i64)it * (P.d + 1) + x] = r; P.rks[x] = r; P.qsnap[(it & 1) * (P.d + 1) + x] = r;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < P.P; v += blockDim.x) {          // one thread per virtual rank: independent L2 round trips
        atomicAdd(&s_ne, (unsigned long long)st[v].neval);
        st[v].pivotmax_prev = st[v].pivotmax;            // dmrgg.f90:961
        st[v].pivotmax = -1.0; st[v].pivotmin = -1.0;    // dmrgg.f90:326-327 of the next sweep
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    P.slog[it].neval = (i64)s_ne; P.slog[it].amax = s_amax; P.slog[it].pivotmax = s_pmax; P.slog[it].pivotmin = s_pmin;
    P.slog[it].t_ns = globaltimer_ns() - P.ctrl->t0_ns; P.slog[it].valid = 1; P.slog[it].pad = 0;
    P.ctrl->nsweeps = it;
    int ready = 0;
    if (maxrank > 0) ready = (it + 1 >= maxrank);
    if (P.ctrl->has_accuracy) {
        if (s_pmax <= P.ctrl->accuracy * s_amax) P.ctrl->strike += 1; else P.ctrl->strike = 0;
        ready = ready || (P.ctrl->strike >= 3);
    }
    if (P.ctrl->error) ready = 1;
    P.ctrl->it = it + 1;
    __threadfence();
    P.ctrl->ready = ready;
}

// ----------------------------------------------------------------------------
// Post-sweep exchange of a single process in ONE kernel (k_exchange_corner + k_exchange_extend_w + k_sweep_log).
// Boundary b, shared core c = own[b+1]: the LEFT chain of mode index x (row(c)(:, x, rc), d2_luar with inv(c-1)) and the
// RIGHT chain of x (col(c)(rc1, x, :), d2_lual with inv(c)) each need exactly one element of the corner fiber, the one
// at x -- so the warp that owns x evaluates it itself and no CTA waits for another one (dmrgg.f90:872-958,
// dmrggmp.f90:572-629).  grid (ceil(2 nmax / warps per CTA), boundaries); the last CTA closes the sweep (close_maxrank > 0).
// dynamic smem: A[auxsm] | XF[d] WF[d] | TL[Rmax^2] | TR[Rmax^2 + Rmax]
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_exchange_fused(DevPlan P, int close_maxrank) {
    tl_stamp(P, 9);
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    __shared__ int s_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int b = blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    const int nc = P.n[c];
    const bool corner = (rc1 > rc1s) && (rc > rcs), left = rc > rcs, right = rc1 > rc1s;
    if (left || right) {
        const double* A = stage_aux<KIND>(P, smem);
        double* XF = smem + P.auxsm; double* WF = XF + P.d;
        double* TL = WF + P.d; double* TR = TL + (i64)P.Rmax * P.Rmax; double* DI = TR + (i64)P.Rmax * P.Rmax;
        double* argc = P.arg + P.coreOff[c];
        const bool hasw = (P.kind == KIND_ISING);
        const int nwoff = P.n[1];
        if (corner) {
            const int* Lt = P.Lidx + P.offL[c - 1]; const int* Rt = P.Ridx + P.offR[c];
            for (int pos = threadIdx.x; pos < P.d - 1; pos += blockDim.x) {
                const int idx = (pos < c - 1) ? Lt[(i64)pos * P.Rmax + (rc1 - 1)] : Rt[(i64)(pos - (c - 1)) * P.Rmax + (rc - 1)];
                XF[pos] = P.par[idx - 1]; WF[pos] = hasw ? P.par[nwoff + idx - 1] : 0.0;
            }
        }
        if (left) stage_luar(P.inv + (i64)(c - 1) * P.Rmax * P.Rmax, rc1, TL);
        if (right) stage_lual(P.inv + (i64)c * P.Rmax * P.Rmax, rc, TR, DI);
        __syncthreads();
        const int job = blockIdx.x * nw + wid;                      // one warp per (mode index, side): both sides evaluate the
        const int x = job >> 1, side = job & 1;                     // corner element (cheap), side 0 stores and counts it
        Partial best = amax_init();
        if (x < nc) {
            double f = 0.0;
            if (corner) {                                           // lane 0 evaluates and stores, the warp gets the value
                if (lane == 0) {
                    StagedVals sv;
                    sv.XL = XF; sv.WL = WF; sv.nl = c - 1; sv.rl = 1; sv.i = 1;
                    sv.xj = P.par[x]; sv.wj = hasw ? P.par[nwoff + x] : 0.0;
                    sv.hask = 0; sv.xk = 0.0; sv.wk = 0.0;
                    sv.XR = XF + (c - 1); sv.WR = WF + (c - 1); sv.rr = 1; sv.q = 1;
                    f = eval_point_wide<KIND>(P, sv, A);
                    if (side == 0) { argc[(rc1 - 1) + (i64)P.Rmax * (x + (i64)nc * (rc - 1))] = f; amax_take(best, f, x); }
                }
                f = __shfl_sync(FULLMASK, f, 0);
            }
            if (left && side == 0) {
                // LEFT receiver (virtual rank b): row(c)(:, x, rc) = d2_luar(rc1, inv(c-1)) of arg(c)(:, x, rc)
                const double* src = argc + (i64)P.Rmax * (x + (i64)nc * (rc - 1));
                double* dst = P.rowT + P.coreOff[c] + (i64)nc * (rc - 1) + x;
                const i64 de = (i64)nc * P.Rmax;
                double y[MAXRPL];
#pragma unroll
                for (int u = 0; u < MAXRPL; ++u) {
                    const int sidx = lane + 32 * u;
                    y[u] = (sidx < rc1) ? ((corner && sidx == rc1 - 1) ? f : src[sidx]) : 0.0;
                }
                warp_luar(y, rc1, GSm{TL, rc1});
#pragma unroll
                for (int u = 0; u < MAXRPL; ++u) { const int sidx = lane + 32 * u; if (sidx < rc1) dst[sidx * de] = y[u]; }
            }
            if (right && side == 1) {
                // RIGHT receiver (virtual rank b+1): col(c)(rc1, x, :) = d2_lual(rc, inv(c)) of arg(c)(rc1, x, :)
                const i64 se = (i64)P.Rmax * nc;
                const double* src = argc + (rc1 - 1) + (i64)P.Rmax * x;
                double* dst = P.col + P.coreOff[c] + (rc1 - 1) + (i64)P.Rmax * x;
                double y[MAXRPL];
#pragma unroll
                for (int u = 0; u < MAXRPL; ++u) {
                    const int cc = lane + 32 * u;
                    y[u] = (cc < rc) ? ((corner && cc == rc - 1) ? f : src[cc * se]) : 0.0;
                }
                warp_lual(y, rc, GSm{TR, rc}, DSm{DI});
#pragma unroll
                for (int u = 0; u < MAXRPL; ++u) { const int cc = lane + 32 * u; if (cc < rc) dst[cc * se] = y[u]; }
            }
        }
        if (corner) {
            best = amax_block(best, shp);
            if (threadIdx.x == 0) {
                // both virtual ranks of the boundary saw the fiber (amax >= 0: the bit pattern orders like the value); each counts it
                for (int v = b; v <= b + 1; ++v) {
                    atomicMax((long long*)&P.st[v].amax, __double_as_longlong(best.absv));
                    if (blockIdx.x == 0) atomicAdd((unsigned long long*)&P.st[v].neval, (unsigned long long)nc);
                }
            }
        }
    }
    if (close_maxrank <= 0) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned tk = atomicAdd(P.btick, 1u);
        s_last = (tk == total - 1);
        if (s_last) { P.btick[0] = 0; __threadfence(); }
    }
    __syncthreads();
    if (s_last) sweep_close_fused(P, close_maxrank);
}

// ----------------------------------------------------------------------------
// all bond visits of one sweep direction for the virtual ranks of this process (pivoting >= 0).
// grid (CS, nv), cluster (CS, 1, 1).  dynamic smem: A[auxsm] | xs[Rmax] | ext[Rmax^2 + Rmax] | stage[stage_max] | ints[4*Rmax + 8]
// fold_allreduce != 0: the last cluster to finish performs the MAX reduction of dmrgg.f90:852-870 (single process only).
// close_maxrank > 0 (single partition only): the cluster also closes the sweep (sweep_close_fused).
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tl_mark(const DevPlan& P, int id) {     // diagnostic: phase stamps of the first cluster
    if (P.tlog && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        int k = atomicAdd(P.tlog_n, 1);
        if (k < P.tlog_cap) { P.tlog[3 * k] = (unsigned long long)id; P.tlog[3 * k + 1] = t; P.tlog[3 * k + 2] = (unsigned long long)clock64(); }
    }
}


// ----------------------------------------------------------------------------
// TMA staging (sm_100a): 1-D bulk copies global -> shared memory completing on an mbarrier (cp.async.bulk, SASS UBLKCP).
// ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Issue the staging of a bond visit (one thread): node / weight vectors of the two free modes and the value tables of the
// left bond pl (pl rows) and the right bond pr (d - pr rows), whole padded rows.  Layout (doubles, all offsets even):
//   NX[NT] NW[NT] NX2[NT] NW2[NT] XL[pl*RT] WL[pl*RT] XR[nr*RT] WR[nr*RT]
__device__ __forceinline__ void stage_bond_tma_issue(const DevPlan& P, double* sm, int pl, int c1, int c2, int pr, unsigned long long* bar) {
    const int RT = P.RT, NT = P.NT, nl = pl, nr = P.d - pr;
    const bool hasw = (P.kind == KIND_ISING);
    const int n1 = (P.n[c1] + 1) & ~1, n2 = (P.n[c2] + 1) & ~1;
    const unsigned bl = (unsigned)(nl * RT * 8), br = (unsigned)(nr * RT * 8);
    const unsigned total = (unsigned)(n1 * 8 + n2 * 8) * (hasw ? 2u : 1u) + (bl + br) * (hasw ? 2u : 1u);
    double* NX = sm; double* NW = NX + NT; double* NX2 = NW + NT; double* NW2 = NX2 + NT;
    double* XL = NW2 + NT; double* WL = XL + (i64)nl * RT; double* XR = WL + (i64)nl * RT; double* WR = XR + (i64)nr * RT;
    const i64 oL = P.offL[pl] / P.Rmax * RT, oR = P.offR[pr] / P.Rmax * RT;
    fence_proxy_async();                       // the staging area was last touched through the generic proxy
    mbar_expect_tx(bar, total);
    tma_load_1d(NX, P.parT, n1 * 8, bar);
    tma_load_1d(NX2, P.parT, n2 * 8, bar);
    if (hasw) { tma_load_1d(NW, P.parT + NT, n1 * 8, bar); tma_load_1d(NW2, P.parT + NT, n2 * 8, bar); }
    if (bl) { tma_load_1d(XL, P.XLg + oL, bl, bar); if (hasw) tma_load_1d(WL, P.WLg + oL, bl, bar); }
    if (br) { tma_load_1d(XR, P.XRg + oR, br, bar); if (hasw) tma_load_1d(WR, P.WRg + oR, br, bar); }
}
__device__ __forceinline__ Stage stage_bond_tma_view(const DevPlan& P, double* sm, int pl, int rl, int pr, int rr) {
    Stage S;
    const int RT = P.RT, NT = P.NT, nl = pl, nr = P.d - pr;
    S.NX = sm; S.NW = sm + NT; S.NX2 = sm + 2 * NT; S.NW2 = sm + 3 * NT;
    S.XL = sm + 4 * NT; S.WL = S.XL + (i64)nl * RT; S.XR = S.WL + (i64)nl * RT; S.WR = S.XR + (i64)nr * RT;
    S.nl = nl; S.rl = RT; S.nr = nr; S.rr = RT; S.hask = 1;      // rl / rr are the row STRIDES of the tables
    (void)rl; (void)rr;
    return S;
}

// ----------------------------------------------------------------------------
// Streaming evaluation of the Ising C integrand (test_crs_ising.f90:186-217, id = 1) at the points of a bond visit.
// The reference runs two independent recurrences over the positions -- the prefix sums w (positions 1..m ascending) and
// the suffix sums v (m..1 descending) -- then f = 2 / (v * w) and the product of the m weights in position order.
// At a bond visit a point is (left pivot i | j | k | right pivot q): the first p-1 prefix steps depend on i only and the
// first m-p-1 suffix steps on q only, so they are taken ONCE per pivot and visit (tables PW*, SV* in shared memory,
// ising_c_prepare) and every evaluation continues them: the same operations on the same operands in the same order
// (bit-identical), read straight from the staged tables -- no per-thread local array, no call, and the dependent chain
// of an evaluation is max(p+1, m-p+1) multiplications instead of 2m.
// ----------------------------------------------------------------------------
struct IsingCTab {
    const double *XL, *WL, *XR, *WR, *NX, *NW, *NX2, *NW2;
    const double *PWK, *PWW, *SVK, *SVV;
    int nl, rl, nr, rr;
};
// pre: shared double[4 * Rmax]; every thread of the CTA calls it (ends with a block barrier)
__device__ __forceinline__ IsingCTab ising_c_prepare(const Stage& S, int r0, int r2, double* pre, int Rmax) {
    IsingCTab T;
    T.XL = S.XL; T.WL = S.WL; T.XR = S.XR; T.WR = S.WR; T.NX = S.NX; T.NW = S.NW; T.NX2 = S.NX2; T.NW2 = S.NW2;
    T.nl = S.nl; T.rl = S.rl; T.nr = S.nr; T.rr = S.rr;
    double* PWK = pre; double* PWW = pre + Rmax; double* SVK = pre + 2 * Rmax; double* SVV = pre + 3 * Rmax;
    for (int t = threadIdx.x; t < r0 + r2; t += blockDim.x) {
        if (t < r0) {
            double wk = 1.0, w = 1.0;
            const double* xl = S.XL + t;
            for (int pos = 0; pos < S.nl; ++pos) { wk = wk * xl[pos * S.rl]; w = w + wk; }
            PWK[t] = wk; PWW[t] = w;
        } else {
            const int q = t - r0;
            double vk = 1.0, vv = 1.0;
            const double* xr = S.XR + q;
            for (int pos = S.nr - 1; pos >= 0; --pos) { vk = vk * xr[pos * S.rr]; vv = vv + vk; }
            SVK[q] = vk; SVV[q] = vv;
        }
    }
    T.PWK = PWK; T.PWW = PWW; T.SVK = SVK; T.SVV = SVV;
    __syncthreads();
    return T;
}
__device__ __forceinline__ double ising_c_eval(const IsingCTab& T, int i, int j, int k, int q) {     // 1-based
    const double xj = T.NX[j - 1], xk = T.NX2[k - 1];
    const double* xl = T.XL + (i - 1); const double* xr = T.XR + (q - 1);
    double wk = T.PWK[i - 1], w = T.PWW[i - 1];
    double vk = T.SVK[q - 1], vv = T.SVV[q - 1];
    wk = wk * xj; w = w + wk;
    vk = vk * xk; vv = vv + vk;
    wk = wk * xk; w = w + wk;
    vk = vk * xj; vv = vv + vk;
#pragma unroll 4
    for (int t = 0; t < T.nr; ++t) { wk = wk * xr[t * T.rr]; w = w + wk; }
#pragma unroll 4
    for (int t = T.nl - 1; t >= 0; --t) { vk = vk * xl[t * T.rl]; vv = vv + vk; }
    const double b = 1.0 / (vv * w);
    double f = 2 * b;
    const double* wl = T.WL + (i - 1); const double* wr = T.WR + (q - 1);
#pragma unroll 4
    for (int t = 0; t < T.nl; ++t) f = f * wl[t * T.rl];
    f = f * T.NW[j - 1];
    f = f * T.NW2[k - 1];
#pragma unroll 4
    for (int t = 0; t < T.nr; ++t) f = f * wr[t * T.rr];
    return f;
}
// the same for a point with ONE free mode (the corner fiber of the exchange, dmrgg.f90:925-937): fixed left part XF[0..nl),
// mode index j, fixed right part XF[nl..nl+nr); W* the weights.  pw / sv: prefix state after the left part, suffix state
// after the right part (computed once per fiber by the caller with ising_c_fixed).
__device__ __forceinline__ void ising_c_fixed(const double* XF, int nl, int nr, double* out4) {
    double wk = 1.0, w = 1.0, vk = 1.0, vv = 1.0;
    for (int t = 0; t < nl; ++t) { wk = wk * XF[t]; w = w + wk; }
    for (int t = nr - 1; t >= 0; --t) { vk = vk * XF[nl + t]; vv = vv + vk; }
    out4[0] = wk; out4[1] = w; out4[2] = vk; out4[3] = vv;
}
__device__ __forceinline__ double ising_c_eval1(const double* XF, const double* WF, int nl, int nr, const double* st4, double xj, double wj) {
    double wk = st4[0], w = st4[1], vk = st4[2], vv = st4[3];
    wk = wk * xj; w = w + wk;
    vk = vk * xj; vv = vv + vk;
#pragma unroll 4
    for (int t = 0; t < nr; ++t) { wk = wk * XF[nl + t]; w = w + wk; }
#pragma unroll 4
    for (int t = nl - 1; t >= 0; --t) { vk = vk * XF[t]; vv = vv + vk; }
    const double b = 1.0 / (vv * w);
    double f = 2 * b;
#pragma unroll 4
    for (int t = 0; t < nl; ++t) f = f * WF[t];
    f = f * wj;
#pragma unroll 4
    for (int t = 0; t < nr; ++t) f = f * WF[nl + t];
    return f;
}
// residuals with ALL factor values of the first 32 terms in flight before the evaluation starts (one L2 round trip per
// element instead of one per batch); same operations in the same order as resid_axpy / resid_dot / resid_ddot2
constexpr int RPF = 32;
struct PrefF { double a[RPF]; };
__device__ __forceinline__ void pref_loadf(PrefF& pf, const double* base, i64 stride, int r) {
#pragma unroll
    for (int u = 0; u < RPF; ++u) pf.a[u] = LDF(base + min(u, r - 1) * stride);   // clamped, never predicated
}
__device__ __forceinline__ double resid_axpy_pff(double f, const PrefF& pf, const double* base, i64 stride, const double* xs, int r) {
    double res = f;
#pragma unroll
    for (int u = 0; u < RPF; ++u) if (u < r) res = res + (-xs[u]) * pf.a[u];
    for (int s0 = RPF; s0 < r; s0 += RP) {
        double a[RP];
#pragma unroll
        for (int u = 0; u < RP; ++u) a[u] = LDF(base + min(s0 + u, r - 1) * stride);
#pragma unroll
        for (int u = 0; u < RP; ++u) if (s0 + u < r) res = res + (-xs[s0 + u]) * a[u];
    }
    return res;
}
__device__ __forceinline__ double resid_dot_pff(double f, const PrefF& pf, const double* base, i64 stride, const double* xs, int r) {
    double t = 0.0;
#pragma unroll
    for (int u = 0; u < RPF; ++u) if (u < r) t = t + pf.a[u] * xs[u];
    for (int s0 = RPF; s0 < r; s0 += RP) {
        double a[RP];
#pragma unroll
        for (int u = 0; u < RP; ++u) a[u] = LDF(base + min(s0 + u, r - 1) * stride);
#pragma unroll
        for (int u = 0; u < RP; ++u) if (s0 + u < r) t = t + a[u] * xs[s0 + u];
    }
    return f + (-t);
}

// Everything a cluster needs to run the bond visits of one sweep: the cluster's shared exchange area, the carved-up dynamic
// shared memory, and the partition's geometry.  rkL / rkR are the sweep-start ranks of the two foreign bonds next to the
// partition (lo-1 and hi): the per-sweep kernel reads them from the snapshot P.rks, the persistent kernel (ttc_sweep.cuh)
// tracks them itself.
struct VisitCtx {
    VisitShared* sh;
    const double* A; double* xs; double* pre; double* ext; double* stg; int* ibuf;   // pre: 4 Rmax doubles (ising_c_prepare)
    int v, lo, hi, rkL, rkR;
    int phase;                 // parity counter of cluster_fold
    int upd_first, upd_last;   // out: was the partition's first / last bond updated in this sweep (uniform over the cluster)
    unsigned long long* bar; unsigned bar_parity;     // TMA staging (persistent kernel): CTA-local mbarrier and its phase
};
// Hook called by every thread right before the accept test of the sweep's FIRST visit: the persistent kernel closes the
// PREVIOUS sweep there (MAX allreduce, record, exit test), so the all-to-all wait for the other partitions' records overlaps
// the lottery and the fibers of this visit, which have no side effects outside the cluster.  Returns true to abandon the sweep.
struct NoVisitHook { __device__ __forceinline__ bool operator()() const { return false; } };
template <int KIND, bool TMA = false, bool WIDE = false, class Hook = NoVisitHook>
__device__ __forceinline__ bool visit_list(const DevPlan& P, cg::cluster_group& cl, VisitCtx& C, int it, int dir, double small_element, double small_pivot,
                                           Hook hook = Hook()) {
    VisitShared& sh = *C.sh;
    const int crank = (int)cl.block_rank(), cs = (int)cl.num_blocks();
    const int v = C.v, lo = C.lo, hi = C.hi, nb = hi - lo;
    const int gtid = crank * blockDim.x + threadIdx.x, gthreads = cs * blockDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double* A = C.A;
    double* xs = C.xs; double* ext = C.ext; double* stg = C.stg; int* ibuf = C.ibuf;
    int& phase = C.phase;
    double* fa_c = P.acol1 + (i64)v * P.Rmax * P.nmax; double* fb_c = P.bcol1 + (i64)v * P.Rmax * P.nmax;
    double* fa_r = P.arow1 + (i64)v * P.Rmax * P.nmax; double* fb_r = P.brow1 + (i64)v * P.Rmax * P.nmax;
    C.upd_first = 0; C.upd_last = 0;

    for (int pp = 1; pp <= nb; ++pp) {
        const int p = (dir == 1) ? lo + pp - 1 : hi - pp;
        if (threadIdx.x == 0) {
            sh.r0 = (p - 1 >= lo) ? LDF(P.rk + p - 1) : C.rkL;
            sh.r1 = LDF(P.rk + p);
            sh.r2 = (p + 1 <= hi - 1) ? LDF(P.rk + p + 1) : C.rkR;
        }
        __syncthreads();
        const int r0 = sh.r0, r1 = sh.r1, r2 = sh.r2, n1 = P.n[p], n2 = P.n[p + 1];
        Stage S;
        if (TMA) {
            // the TMA fills the staging area (the previous visit's readers are past the cluster barrier that ended it)
            if (threadIdx.x == 0) stage_bond_tma_issue(P, stg, p - 1, p, p + 1, p + 1, C.bar);
            S = stage_bond_tma_view(P, stg, p - 1, r0, p + 1, r2);
        } else {
            S = stage_bond_cg(P, stg, p - 1, r0, p, p + 1, p + 1, r2);
        }
        constexpr bool fastc = (KIND == KIND_ISINGC);                      // streaming evaluation of Ising C
        IsingCTab TC;
        tl_mark(P, 41);
        const double* colp = P.col + P.coreOff[p];
        const double* rowp = P.rowT + P.coreOff[p + 1];
        const i64 cs_ = (i64)P.Rmax * n1;        // stride of s in col(i,j,s)
        const i64 rs_ = (i64)n2 * P.Rmax;        // stride of s in rowT(s,k,q)
        const int ccount = r0 * n1, rcount = n2 * r2;

        // ---- lottery candidates (dmrgg.f90:425-490)
        {
            const int nlot = r0 + n1 + n2 + r2;
            int* tmp = ibuf; int* zc = ibuf + 2 * P.Rmax; int* zr = ibuf + 3 * P.Rmax;
            const int* vip_p = P.vip + (i64)p * P.Rmax * 4;
            const int m = ccount, n = rcount;
            lot_zeros2(vip_p, r1, r0, n2, tmp, zc, zr, sh.nz);
            if (TMA) { mbar_wait(C.bar, C.bar_parity); C.bar_parity ^= 1u; }     // the staged tables have landed (copied beside the zero-cell setup)
            if (fastc) TC = ising_c_prepare(S, r0, r2, C.pre, P.Rmax);
            tl_mark(P, 42);
            const unsigned long long k0 = sh.S.rng_k, seed = P.ctrl->seed;
            const bool packed = m < (1 << 20) && n < (1 << 20) && nlot < (1 << 22);
            Partial bres = amax_init();
            double braw = -1.0;                  // largest |f| (NaN never wins: fmax drops it, like the strict '>' of idamax)
            // a lane pair shares one candidate: the even lane draws its column cell, the odd lane its row cell (the two
            // bisections are the long serial part of a candidate), then the even lane evaluates
            for (int x0 = 0; x0 < nlot; x0 += gthreads / 2) {
                const int x = x0 + (gtid >> 1);
                const int side = gtid & 1;
                const bool live = x < nlot;
                int cell = 1;
                if (live) {
                    const double uu = stream_uniform(seed, v, k0 + (unsigned long long)(side ? nlot + x : x));
                    cell = side ? lot_draw_fast(n - sh.nz[1], n, zr, sh.nz[1], uu) : lot_draw_fast(m - sh.nz[0], m, zc, sh.nz[0], uu);
                }
                if (fastc) {
                    // streaming path: BOTH lanes of the pair know the candidate and split the factor loads of the residual
                    // (each the halves [0,16) / [16,32) of every 32 terms), so the 2 r values are in flight in one L2 round
                    // trip; the even lane then adds the products in the reference order, its partner's arriving by shuffle
                    const int other = __shfl_xor_sync(FULLMASK, cell, 1);
                    const int c = side ? other : cell, w = side ? cell : other;
                    const int i = (c - 1) % r0 + 1, j = (c - 1) / r0 + 1, k = (w - 1) % n2 + 1, q = (w - 1) / n2 + 1;
                    const double* cp = colp + (i - 1) + (i64)P.Rmax * (j - 1);
                    const double* rp = rowp + (k - 1) + (i64)n2 * (q - 1);
                    double la[16], lb[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) { const int sx = min(side * 16 + u, r1 - 1); la[u] = LDF(cp + sx * cs_); lb[u] = LDF(rp + sx * rs_); }
                    tl_mark(P, 55);
                    const double f = ising_c_eval(TC, i, j, k, q);
                    tl_mark(P, 56);
                    double tsum = 0.0;
                    for (int s0 = 0; s0 < r1; s0 += 32) {
                        if (s0 > 0) {
#pragma unroll
                            for (int u = 0; u < 16; ++u) { const int sx = min(s0 + side * 16 + u, r1 - 1); la[u] = LDF(cp + sx * cs_); lb[u] = LDF(rp + sx * rs_); }
                        }
                        double pa[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) pa[u] = la[u] * lb[u];
#pragma unroll
                        for (int u = 0; u < 16; ++u) if (s0 + u < r1) tsum = tsum + pa[u];
#pragma unroll
                        for (int u = 0; u < 16; ++u) { const double po = __shfl_down_sync(FULLMASK, pa[u], 1); if (s0 + 16 + u < r1) tsum = tsum + po; }
                    }
                    const double res = f - tsum;
                    if (live && !side) {
                        braw = fmax(braw, fabs(f));
                        amax_take(bres, res, packed ? (((i64)x << 40) | ((i64)c << 20) | (i64)w) : (i64)x);
                    }
                    continue;
                }
                const int w = __shfl_down_sync(FULLMASK, cell, 1);
                tl_mark(P, 55);
                if (!live || side) continue;
                const int c = cell;
                const int i = (c - 1) % r0 + 1, j = (c - 1) / r0 + 1, k = (w - 1) % n2 + 1, q = (w - 1) / n2 + 1;
                StagedVals sv = S.point(i, j, k, q);
                const double f = WIDE ? eval_point_wide<KIND>(P, sv, A) : eval_point<KIND>(P, sv, A);
                tl_mark(P, 56);
                const double res = resid_ddot2_cg(f, colp + (i - 1) + (i64)P.Rmax * (j - 1), cs_, rowp + (k - 1) + (i64)n2 * (q - 1), rs_, r1);
                braw = fmax(braw, fabs(f));
                // the winner's cells ride along below the draw number (the draw number stays the major key of the
                // first-index tie-break), so nobody has to redraw them after the fold
                amax_take(bres, res, packed ? (((i64)x << 40) | ((i64)c << 20) | (i64)w) : (i64)x);
            }
            tl_mark(P, 43);
            cluster_fold(cl, sh, phase, braw, bres);
            tl_mark(P, 44);
            if (threadIdx.x == 0) {
                VState& St = sh.S;
                St.amax = fmax(St.amax, braw);
                int c, w;
                if (packed) { c = (int)((bres.idx >> 20) & 0xfffff); w = (int)(bres.idx & 0xfffff); }
                else {                               // the winner's cell is a pure function of its draw number
                    const int x = (int)bres.idx;
                    const double uc = stream_uniform(seed, v, k0 + (unsigned long long)x);
                    const double ur = stream_uniform(seed, v, k0 + (unsigned long long)(nlot + x));
                    c = lot_draw_fast(m - sh.nz[0], m, zc, sh.nz[0], uc);
                    w = lot_draw_fast(n - sh.nz[1], n, zr, sh.nz[1], ur);
                }
                St.ii = (c - 1) % r0 + 1; St.jj = (c - 1) / r0 + 1; St.kk = (w - 1) % n2 + 1; St.qq = (w - 1) / n2 + 1;
                St.pivot = bres.val;
                St.done = 0; St.havecol = 0; St.haverow = 0; St.crs = 0; St.upd = 0;
                St.neval += nlot;
                St.rng_k += 2ULL * (unsigned long long)nlot;
            }
            __syncthreads();
        }

        // ---- cross fibers: piv = 0 one column + one row (dmrgg.f90:492-513), piv >= 1 the rook loop (:515-582)
        const int nfib = (P.piv == 0) ? 2 : 2 * P.piv;
        int isrow = (P.piv == 0) ? 0 : ((dir == 2) ? 1 : 0);
        for (int c = 0; c < nfib; ++c, isrow ^= 1) {
            if (sh.S.done) break;                                   // uniform over the cluster
            const int ii = sh.S.ii, jj = sh.S.jj, kk = sh.S.kk, qq = sh.S.qq;
            const int count = isrow ? rcount : ccount;
            double* fa = isrow ? fa_r : fa_c;
            double* fb = isrow ? fb_r : fb_c;
            Partial bres = amax_init();
            double braw = -1.0;
            // streaming path: the factor values of this thread's first element are requested BEFORE the staging of xs, so the
            // two L2 round trips of a fiber step overlap
            int e = gtid, a1 = 1, a2 = 1;
            const i64 fstride = isrow ? rs_ : cs_;
            const double* base = colp;
            PrefF pff;
            if (fastc && e < count) {
                if (!isrow) { a2 = e / r0 + 1; a1 = e % r0 + 1; base = colp + (a1 - 1) + (i64)P.Rmax * (a2 - 1); }
                else { a2 = e / n2 + 1; a1 = e % n2 + 1; base = rowp + (a1 - 1) + (i64)n2 * (a2 - 1); }
                pref_loadf(pff, base, fstride, r1);
            }
            __syncthreads();                                        // xs of the previous fiber no longer read
            for (int s = threadIdx.x; s < r1; s += blockDim.x)
                xs[s] = isrow ? LDF(colp + (ii - 1) + (i64)P.Rmax * (jj - 1) + s * cs_) : LDF(rowp + (kk - 1) + (i64)n2 * (qq - 1) + s * rs_);
            __syncthreads();
            tl_mark(P, 50);
            if (fastc) {
                while (e < count) {
                    double f, res;
                    if (!isrow) { f = ising_c_eval(TC, a1, a2, kk, qq); res = resid_axpy_pff(f, pff, base, fstride, xs, r1); }
                    else { f = ising_c_eval(TC, ii, jj, a1, a2); res = resid_dot_pff(f, pff, base, fstride, xs, r1); }
                    fa[e] = f;
                    fb[e] = res;
                    braw = fmax(braw, fabs(f));
                    amax_take(bres, res, e);
                    e += gthreads;
                    if (e < count) {
                        if (!isrow) { a2 = e / r0 + 1; a1 = e % r0 + 1; base = colp + (a1 - 1) + (i64)P.Rmax * (a2 - 1); }
                        else { a2 = e / n2 + 1; a1 = e % n2 + 1; base = rowp + (a1 - 1) + (i64)n2 * (a2 - 1); }
                        pref_loadf(pff, base, fstride, r1);
                    }
                }
            } else {
            for (e = gtid; e < count; e += gthreads) {
                double f, res;
                if (KIND == KIND_MVN) {
                    // an MVN evaluation is a 3 d^2-long dependent chain (tens of microseconds): nothing to hide behind it, and the
                    // register-resident evaluation wants every register -> factor values are loaded AFTER it, in batches
                    if (!isrow) {
                        const int j = e / r0 + 1, i = e % r0 + 1;
                        StagedVals sv = S.point(i, j, kk, qq);
                        f = WIDE ? eval_point_wide<KIND>(P, sv, A) : eval_point<KIND>(P, sv, A);
                        res = resid_axpy_cg(f, colp + (i - 1) + (i64)P.Rmax * (j - 1), cs_, xs, r1);
                    } else {
                        const int q = e / n2 + 1, k = e % n2 + 1;
                        StagedVals sv = S.point(ii, jj, k, q);
                        f = WIDE ? eval_point_wide<KIND>(P, sv, A) : eval_point<KIND>(P, sv, A);
                        res = resid_dot_cg(f, rowp + (k - 1) + (i64)n2 * (q - 1), rs_, xs, r1);
                    }
                } else {
                Pref pf;
                if (!isrow) {
                    const int j = e / r0 + 1, i = e % r0 + 1;
                    const double* base = colp + (i - 1) + (i64)P.Rmax * (j - 1);
                    pref_load(pf, base, cs_, r1);
                    StagedVals sv = S.point(i, j, kk, qq);
                    f = WIDE ? eval_point_wide<KIND>(P, sv, A) : eval_point<KIND>(P, sv, A);
                    res = resid_axpy_pf(f, pf, base, cs_, xs, r1);
                } else {
                    const int q = e / n2 + 1, k = e % n2 + 1;
                    const double* base = rowp + (k - 1) + (i64)n2 * (q - 1);
                    pref_load(pf, base, rs_, r1);
                    StagedVals sv = S.point(ii, jj, k, q);
                    f = WIDE ? eval_point_wide<KIND>(P, sv, A) : eval_point<KIND>(P, sv, A);
                    res = resid_dot_pf(f, pf, base, rs_, xs, r1);
                }
                }
                fa[e] = f;
                fb[e] = res;
                braw = fmax(braw, fabs(f));
                amax_take(bres, res, e);
            }
            }
            tl_mark(P, 45);
            cluster_fold(cl, sh, phase, braw, bres);
            tl_mark(P, 46);
            if (threadIdx.x == 0) {                                 // dmrgg.f90:527-547, 560-580
                VState& St = sh.S;
                St.neval += count;
                if (P.piv == 0) {
                    St.havecol = 1; St.haverow = 1;
                    if (c == 1) St.done = 1;
                } else {
                    St.amax = fmax(St.amax, braw);
                    if (isrow) St.haverow = 1; else St.havecol = 1;
                    St.crs += 1;
                    int done = St.havecol && St.haverow && (St.crs >= 2 * P.piv);
                    if (!done) {
                        const int e = (int)bres.idx;
                        if (!isrow) {
                            const int j = e / r0 + 1, i = e % r0 + 1;
                            done = St.havecol && St.haverow && (i == St.ii && j == St.jj);
                            St.ii = i; St.jj = j;
                        } else {
                            const int q = e / n2 + 1, k = e % n2 + 1;
                            done = St.havecol && St.haverow && (k == St.kk && q == St.qq);
                            St.kk = k; St.qq = q;
                        }
                        St.pivot = bres.val;
                    }
                    St.done = done;
                }
            }
            __syncthreads();
        }

        tl_mark(P, 47);
        if (pp == 1 && hook()) return true;                         // (uniform over the cluster)
        // ---- accept test and index-set update (dmrgg.f90:598-660); every CTA takes the same decision
        const int ii = sh.S.ii, jj = sh.S.jj, kk = sh.S.kk, qq = sh.S.qq;
        const double pivot = sh.S.pivot;
        const double ap = fabs(pivot);
        int upd = (ap > small_element * sh.S.amax) && (ap > small_pivot * sh.S.pivotmax_prev);
        if (upd && r1 >= P.Rmax) { upd = 0; if (crank == 0 && threadIdx.x == 0) P.ctrl->error = 1; }
        __syncthreads();
        if (threadIdx.x == 0) {
            VState& St = sh.S;
            St.upd = upd;
            if (upd) {
                St.pivotmax = (St.pivotmax < 0.0) ? ap : fmax(St.pivotmax, ap);
                St.pivotmin = (St.pivotmin < 0.0) ? ap : fmin(St.pivotmin, ap);
            }
            if (crank == 0) {
                VisitOut& O = P.vlog[((i64)(it - 1) * P.maxnb + (pp - 1)) * P.P + v];
                O.active = 1; O.upd = upd; O.bond = p; O.ii = ii; O.jj = jj; O.kk = kk; O.qq = qq; O.pivot = pivot;
            }
        }
        if (upd) {
            const int t = r1;                                       // 0-based slot of the new pivot
            // the three appends are independent: three CTAs of the cluster take one each (all of them CTA 0 in a one-CTA cluster)
            const bool vt = P.XLg != nullptr, hasw_ = (P.kind == KIND_ISING);
            const int nwoff_ = P.n[1];
            // node value / weight of mode index idx: from the staged vectors when it lies inside them (always for equal mode sizes)
            auto nodev = [&](int idx) { return idx <= n1 ? S.NX[idx - 1] : P.par[idx - 1]; };
            auto nodew = [&](int idx) { return idx <= n1 ? S.NW[idx - 1] : P.par[nwoff_ + idx - 1]; };
            if (crank == 0) {
                if (threadIdx.x == 0) { int* vp = P.vip + ((i64)p * P.Rmax + t) * 4; vp[0] = ii; vp[1] = jj; vp[2] = kk; vp[3] = qq; }
                int* Lp = P.Lidx + P.offL[p];
                const int* Lm = P.Lidx + P.offL[p - 1];
                const i64 oLp = P.offL[p] / P.Rmax * P.RT;
                for (int pos = threadIdx.x; pos < p; pos += blockDim.x) {
                    const int idx = (pos < p - 1) ? LDF(Lm + (i64)pos * P.Rmax + (ii - 1)) : jj;
                    Lp[(i64)pos * P.Rmax + t] = idx;
                    if (vt) { P.XLg[oLp + (i64)pos * P.RT + t] = nodev(idx); if (hasw_) P.WLg[oLp + (i64)pos * P.RT + t] = nodew(idx); }
                }
                if (vt) fence_proxy_async();          // the TMA of later visits reads these tables through the async proxy
            }
            if (crank == 1 % cs) {
                int* Rp = P.Ridx + P.offR[p];
                const int* Rn = P.Ridx + P.offR[p + 1];
                const i64 oRp = P.offR[p] / P.Rmax * P.RT;
                for (int pos = threadIdx.x; pos < P.d - p; pos += blockDim.x) {
                    const int idx = (pos == 0) ? kk : LDF(Rn + (i64)(pos - 1) * P.Rmax + (qq - 1));
                    Rp[(i64)pos * P.Rmax + t] = idx;
                    if (vt) { P.XRg[oRp + (i64)pos * P.RT + t] = nodev(idx); if (hasw_) P.WRg[oRp + (i64)pos * P.RT + t] = nodew(idx); }
                }
                if (vt) fence_proxy_async();
            }
            if (crank == 2 % cs) {
                // packed LU: [ col(ii,jj,1:r) | row(1:r,kk,qq) | pivot ]
                double* g = P.inv + (i64)p * P.Rmax * P.Rmax;
                for (int s = threadIdx.x; s < 2 * r1; s += blockDim.x) {
                    const int s1 = s < r1 ? s : s - r1;
                    g[(i64)r1 * r1 + s] = s < r1 ? LDF(colp + (ii - 1) + (i64)P.Rmax * (jj - 1) + s1 * cs_) : LDF(rowp + (kk - 1) + (i64)n2 * (qq - 1) + s1 * rs_);
                }
                if (threadIdx.x == 0) g[(i64)(r1 + 1) * (r1 + 1) - 1] = pivot;
            }
            // neighbour factors (dmrgg.f90:715-749): warps of the whole cluster run the chains as wavefronts
            if (p > lo && r0 >= 1) {
                stage_luar_cg(P.inv + (i64)(p - 1) * P.Rmax * P.Rmax, r0, ext);
                __syncthreads();
                double* dst = P.rowT + P.coreOff[p] + (i64)n1 * t;
                const i64 de = (i64)n1 * P.Rmax;
                for (int x = crank * nw + wid; x < n1; x += cs * nw) {
                    double y[MAXRPL];
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { int sidx = lane + 32 * u; y[u] = (sidx < r0) ? LDF(fa_c + (i64)x * r0 + sidx) : 0.0; }
                    warp_luar(y, r0, GSm{ext, r0});
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { int sidx = lane + 32 * u; if (sidx < r0) dst[x + sidx * de] = y[u]; }
                }
                __syncthreads();
            }
            if (p < hi - 1 && r2 >= 1) {
                double* di = ext + r2 * r2;
                stage_lual_cg(P.inv + (i64)(p + 1) * P.Rmax * P.Rmax, r2, ext, di);
                __syncthreads();
                double* dst = P.col + P.coreOff[p + 1] + t;
                const i64 de = (i64)P.Rmax * n2;
                for (int x = crank * nw + wid; x < n2; x += cs * nw) {
                    double y[MAXRPL];
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { int c = lane + 32 * u; y[u] = (c < r2) ? LDF(fa_r + x + (i64)c * n2) : 0.0; }
                    warp_lual(y, r2, GSm{ext, r2}, DSm{di});
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { int c = lane + 32 * u; if (c < r2) dst[(i64)x * P.Rmax + c * de] = y[u]; }
                }
                __syncthreads();
            }
            tl_mark(P, 48);
            // rank-1 append (dmrgg.f90:663-713): the residuals of the last fibers ARE the lual/luar(from=r+1) eliminations
            {
                double* argp = P.arg + P.coreOff[p];
                double* argn = P.arg + P.coreOff[p + 1];
                double* colw = P.col + P.coreOff[p];
                double* roww = P.rowT + P.coreOff[p + 1];
                const double sc = 1.0 / pivot;
                // (the loads of a batch are all in flight before the first store: one L2 round trip per batch, not per element)
                constexpr int AU = 4;
                const int tot = ccount + rcount;
                for (int e0 = gtid; e0 < tot; e0 += AU * gthreads) {
                    double va[AU], vb[AU];
#pragma unroll
                    for (int u = 0; u < AU; ++u) {
                        const int e = e0 + u * gthreads;
                        const bool isc = e < ccount;
                        const int x = isc ? e : e - ccount;
                        va[u] = e < tot ? LDF((isc ? fa_c : fa_r) + x) : 0.0;
                        vb[u] = e < tot ? LDF((isc ? fb_c : fb_r) + x) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < AU; ++u) {
                        const int e = e0 + u * gthreads;
                        if (e >= tot) continue;
                        if (e < ccount) {
                            const int j = e / r0, i = e % r0;
                            const i64 o = i + (i64)P.Rmax * (j + (i64)n1 * t);
                            argp[o] = va[u];
                            colw[o] = sc * vb[u];
                        } else {
                            const int x = e - ccount;
                            const int q = x / n2, k = x % n2;
                            argn[t + (i64)P.Rmax * (k + (i64)n2 * q)] = va[u];
                            roww[k + (i64)n2 * (q + (i64)P.Rmax * t)] = vb[u];
                        }
                    }
                }
            }
            if (crank == 0 && threadIdx.x == 0) P.rk[p] = r1 + 1;
            if (p == lo) C.upd_first = 1;
            if (p == hi - 1) C.upd_last = 1;
        }
        tl_mark(P, 49);
        cl.sync();          // the next visit of this cluster (Gauss-Seidel, dmrgg.f90:329-331) sees everything written above
    }
    if (crank == 0) {
        for (int pp = nb + threadIdx.x + 1; pp <= P.maxnb; pp += blockDim.x) {
            VisitOut& O = P.vlog[((i64)(it - 1) * P.maxnb + (pp - 1)) * P.P + v];
            O.active = 0; O.upd = 0;
        }
    }
    return false;
}

constexpr int VISIT_MAXTHREADS = 256;
// MVN evaluations are one long dependent DADD chain each (3 d^2 operations, one accumulator, mvn_pdf.f90:74-80): what
// hides the latency is resident warps, not registers -> 64 registers per thread.  (Measured on config E: the
// register-resident eval_mvn_reg makes one evaluation 2.3x faster, 133 us instead of 300 us, but needs 255 registers =
// one CTA per SM, and 63 clusters then run in four waves instead of two: 86 ms instead of 59 ms.  It is used where a
// single evaluation is the critical path: the corner fibers of the exchange.)
template <int KIND>
__global__ void __launch_bounds__(VISIT_MAXTHREADS, KIND == KIND_MVN ? 4 : 2) k_visits(DevPlan P, int dir, double small_element, double small_pivot, int fold_allreduce, int close_maxrank) {
    tl_stamp(P, 40);
    if (LDF(&P.ctrl->ready)) return;          // uniform over the grid: written only by k_sweep_log
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ double smem[];
    __shared__ VisitShared sh;
    const int crank = (int)cl.block_rank();
    VisitCtx C;
    C.sh = &sh;
    C.v = P.v0 + blockIdx.y;
    C.lo = P.own[C.v]; C.hi = P.own[C.v + 1];
    C.rkL = P.rks[C.lo - 1]; C.rkR = P.rks[C.hi];
    C.A = stage_aux<KIND>(P, smem);
    C.xs = smem + P.auxsm;
    C.pre = C.xs + P.Rmax;
    C.ext = C.pre + 4 * (i64)P.Rmax;
    C.stg = C.ext + (i64)P.Rmax * P.Rmax + P.Rmax;
    C.ibuf = (int*)(C.stg + P.stage_max);
    C.phase = 0;
    const int v = C.v;
    const int it = P.ctrl->it;
    if (threadIdx.x == 0) sh.S = P.st[v];
    visit_list<KIND>(P, cl, C, it, dir, small_element, small_pivot);
    if (close_maxrank > 0) {
        // ---- single partition: this cluster is the whole sweep; close it here (no exchange, no reduction)
        if (crank == 0 && threadIdx.x == 0) P.st[v] = sh.S;
        __threadfence();
        cl.sync();
        if (crank == 0) {
            if (threadIdx.x == 0) {
                __threadfence();
                const unsigned tk = atomicAdd(P.tickets + P.P, 1u);
                sh.nz[1] = (tk == (unsigned)P.nv - 1);
                if (sh.nz[1]) { P.tickets[P.P] = 0; __threadfence(); }
            }
            __syncthreads();
            if (sh.nz[1]) sweep_close_fused(P, close_maxrank);
        }
        return;
    }
    if (crank == 0) {
        if (threadIdx.x == 0) {
            P.st[v] = sh.S;
            if (fold_allreduce && P.P > 1) {
                __threadfence();
                const unsigned tk = atomicAdd(P.tickets + P.P, 1u);
                if (tk == (unsigned)P.nv - 1) {
                    P.tickets[P.P] = 0;
                    __threadfence();
                    volatile VState* st = P.st;
                    double c1 = st[0].amax, c2 = st[0].pivotmax, c3 = (st[0].pivotmin > 0.0) ? -st[0].pivotmin : -999e9;
                    for (int u = 1; u < P.P; ++u) {
                        c1 = fmax(c1, st[u].amax); c2 = fmax(c2, st[u].pivotmax);
                        const double pm = st[u].pivotmin;
                        c3 = fmax(c3, (pm > 0.0) ? -pm : -999e9);
                    }
                    for (int u = 0; u < P.P; ++u) {
                        st[u].amax = c1; st[u].pivotmax = c2;
                        st[u].pivotmin = (-c3 == 999e9) ? -1.0 : -c3;
                    }
                }
            }
        }
    }
}

}  // namespace ttc
