// =============================================================================
// ttc_sweep.cuh — the WHOLE sweep loop of dtt_dmrgg (dmrgg.f90:314-1020) as ONE persistent kernel.
//
// One thread-block cluster per partition (virtual rank) lives across all sweeps.  Per sweep a cluster
//   1. runs the bond visits of its partition (visit_list, ttc_visit.cuh: lottery -> rook fibers -> accept -> rank-1 append),
//   2. announces them (flag1) and waits for its two NEIGHBOURS only,
//   3. does its half of the post-sweep exchange on both of its boundaries (dmrgg.f90:872-958, dmrggmp.f90:572-629): the
//      left member of a boundary appends the new column of row(c) (d2_luar with inv(c-1)), the right member the new row
//      of col(c) (d2_lual with inv(c)); both evaluate the corner fiber when both adjacent bonds grew,
//   4. publishes its scalars (flag2), waits for everybody's, and takes the MAX allreduce (:852-870), the sweep record and
//      the exit test (:1010-1019) -- every cluster computes the same decision from the same records.
// Nothing returns to the host between sweeps and nothing is re-launched: the launch gaps, the per-launch re-staging of the
// control state and the separate exchange / close kernels of the per-sweep schedule are gone.  The per-sweep quadrature
// values (dmrgg.f90:975-993) do not feed the exit test; they are computed after the loop for all sweeps at once
// (k_quad_chain_all / k_quad_tree_all below) from the rank log: every entry of a contracted, luar'd and lual'd core is a
// fixed operation sequence on data that never changes once written, so the value of sweep s is the chain over the
// leading r_s x r_s blocks of the FINAL contracted cores, bit for bit.
//
// The clusters wait on one another (flags in global memory), so the kernel is launched cooperatively (co-residency
// guaranteed by the driver); with several processes (one per GPU) the flags and the boundary slabs live in the peer
// windows (CUDA IPC) and the same kernel pushes them over NVLink.
// =============================================================================
#pragma once
#include "ttc_visit.cuh"

namespace ttc {

struct SweepRec {                      // what a partition tells the others at the end of a sweep
    double amax1, pivotmax, pivotmin;  // after the bond visits: inputs of the MAX allreduce (dmrgg.f90:852-870)
    double amax2;                      // after the corner fibers (rank 0's goes into the sweep record and the exit test)
    long long neval;                   // after the corner fibers
    int error, pad;
};
struct SweepMail {                     // one per partition in every process's window; records double-buffered by sweep parity
    unsigned long long flag1;          // sequence number of the last sweep whose bond visits (and pushes) are complete
    unsigned long long flag2;          // ... whose exchange is complete (rec valid)
    int upd1[2][2];                    // [parity][first bond, last bond] updated in that sweep (valid with flag1)
    SweepRec rec[2];                   // (valid with flag2)
    unsigned long long pad[16];
};
static_assert(sizeof(SweepMail) == 256, "SweepMail is 256 bytes");

struct SweepLocal {                    // per-CTA replica of the cluster's loop state (advanced identically in every CTA)
    int rkL, rkR;                      // current ranks of the foreign bonds lo-1 and hi
    int updL, updR;                    // did the neighbour's adjacent bond grow in this sweep
    int strike, ready, corner, pending;   // pending: sweep whose close (allreduce, record, exit test) has not been taken yet
    double amax1;
    VState saved;                          // the partition's state when `pending` ended (restored if the next sweep is abandoned)
    double r_amax1[64], r_pmax[64], r_pmin[64], r_amax2[64];
    long long r_neval[64];
    int r_err[64];
};

template <bool SYS>
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
    if (SYS) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
    else asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
template <bool SYS>
__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
    unsigned long long v;
    if (SYS) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <class T>
__device__ __forceinline__ T* peer_ptr(const DevPlan& P, int g, long long off) { return (T*)(P.peer_win[g] + off); }
__device__ __forceinline__ int proc_of(const DevPlan& P, int v) {        // the process that runs partition v (block map)
    int g = 0;
    while (g + 1 < P.nproc && proc_v0(P.P, P.nproc, g + 1) <= v) ++g;
    return g;
}
// bounded spin: a lost partner becomes ctrl->error = 3 (TTC_ERR_COMM) instead of a hang
__device__ __forceinline__ void sweep_wait(const DevPlan& P, const unsigned long long* flag, unsigned long long seq, bool sys) {
    unsigned long long t0 = 0;
    int spins = 0;
    while ((sys ? ld_acquire<true>(flag) : ld_acquire<false>(flag)) < seq) {
        if (++spins < 64) continue;
        spins = 0;
        const unsigned long long t1 = globaltimer_ns();
        if (t0 == 0) t0 = t1;
        else if (t1 - t0 > 4000000000ULL) { P.ctrl->error = 3; break; }
    }
}

// ----------------------------------------------------------------------------
// NXI independent d2_luar / d2_lual recurrences of rank r <= 32 in one warp, interleaved: a recurrence is r dependent steps
// of (broadcast, multiply, add) -- about 50 cycles each and nothing to overlap inside one chain -- so a warp that owns
// several mode indices runs them side by side.  Same operations in the same order as warp_luar / warp_lual per chain.
// ----------------------------------------------------------------------------
// d2_luar / d2_lual of ONE mode index per THREAD, rank <= RB, straight out of the packed LU block staged in shared memory
// (every lane reads the same coefficient: a broadcast).  A recurrence is a chain of dependent steps, but only the LAST term of
// row s waits for y(s-1); the r^2/2 other multiply-adds are independent of it, so with the rows fully unrolled in registers
// the dependent path is 3 operations per row and the rest pipelines -- no shuffles, no idle lanes: 32 mode indices per warp
// instead of one.  Same operations in the same order per row as lr.f90:124-154 (tmp from 0 in ascending u, then y - tmp; resp.
// y + (-g) y(u) in ascending u, then the division by the pivot as a multiplication by its reciprocal, like warp_lual).
// Rows >= r work on zeros and are discarded.  G: packed block (dmrgg.f90:650-660), DI[c] = 1 / pivot(c).
// (The loops run over u OUTSIDE and the rows inside: consecutive instructions then belong to different rows and are
// independent, which is what an in-order pipeline needs; every row still receives its terms in ascending u.)
template <int RB>
__device__ __forceinline__ void thread_luar(double (&y)[RB], const double* G) {
    double tmp[RB];
#pragma unroll
    for (int s = 0; s < RB; ++s) tmp[s] = 0.0;
#pragma unroll
    for (int u = 0; u + 1 < RB; ++u) {
#pragma unroll
        for (int s = u + 1; s < RB; ++s) tmp[s] = tmp[s] + y[u] * G[s * s + u];
        y[u + 1] = y[u + 1] + (-tmp[u + 1]);
    }
}
template <int RB>
__device__ __forceinline__ void thread_lual(double (&y)[RB], const double* G, const double* DI) {
    y[0] = DI[0] * y[0];
#pragma unroll
    for (int u = 0; u + 1 < RB; ++u) {
#pragma unroll
        for (int c = u + 1; c < RB; ++c) y[c] = y[c] + (-G[(c + 1) * (c + 1) - (c + 1) + u]) * y[u];
        y[u + 1] = DI[u + 1] * y[u + 1];
    }
}
// (A loop version with the vector in a private shared-memory column was measured too: 19 us per boundary against 8 us for
// the unrolled registers and 9 us for six interleaved warp wavefronts -- the store -> load round trips of the column
// serialise it.)
// one boundary job for the mode index x of this thread: load the right-hand side (the corner value replaces its last entry),
// solve, store.  src / dst strides in doubles.
template <int RB>
__device__ __forceinline__ void thread_boundary(int side, int r, const double* src, i64 sstride, double* dst, i64 dstride, bool corner, double f,
                                                const double* G, const double* DI) {
    double y[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) y[u] = (u < r) ? ((corner && u == r - 1) ? f : __ldcg(src + u * sstride)) : 0.0;
    if (side == 0) thread_luar<RB>(y, G); else thread_lual<RB>(y, G, DI);
#pragma unroll
    for (int u = 0; u < RB; ++u) if (u < r) dst[u * dstride] = y[u];
}

// ----------------------------------------------------------------------------
// This partition's half of the post-sweep exchange (dmrgg.f90:872-958, dmrggmp.f90:572-629) on its two boundaries.
// Core c is shared by the two members of a boundary; rc1 / rc are the ranks of the bonds c-1 / c AFTER the sweep.
//   side 1 (left boundary, c = lo): this partition is the RIGHT member (owns bond c, inv(c), col(c)); if bond c-1 grew, the
//           new row arg(c)(rc1, x, :) arrived and col(c)(rc1, x, :) = d2_lual(rc, inv(c)) of it is appended (dmrggmp.f90:597-626);
//   side 0 (right boundary, c = hi): the LEFT member (owns bond c-1, inv(c-1), row(c)); if bond c grew, the new slice
//           arg(c)(:, x, rc) arrived and row(c)(:, x, rc) = d2_luar(rc1, inv(c-1)) of it is appended (dmrgg.f90:913-951).
// When both bonds of a boundary grew, the corner fiber arg(c)(rc1, x, rc), x = 1..n(c), is evaluated by BOTH members (each
// counts it, like the reference).  When both boundaries have work the CTAs of the cluster split into two halves; a CTA
// first evaluates the corner values of exactly the mode indices its own warps own (one thread each), then its warps run
// their recurrences interleaved.  Returns the largest |corner value| seen by this thread (-1: none).
// smem: ext = staged LU table (+ diagonal), stg = XF[d] | WF[d] | F[corner values of this CTA].
// ----------------------------------------------------------------------------
struct BoundaryJob { int exists, work, side, c, rc1, rc, corner; };   // exists: the partition has this boundary; work: something arrived
template <int KIND>
__device__ __forceinline__ double exchange_boundaries(const DevPlan& P, cg::cluster_group& cl, const VisitCtx& C, const BoundaryJob& JA, const BoundaryJob& JB,
                                                      int stage_only, int prestaged) {
    if (!JA.work && !JB.work) return -1.0;
    const int crank = (int)cl.block_rank(), cs = (int)cl.num_blocks();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // which boundary this CTA serves, and its position among the CTAs that serve it: a STATIC split (inner partitions: the
    // first half of the cluster serves the left boundary, the rest the right one), so that a CTA can stage its packed LU --
    // this partition's own data -- before the neighbour's flag arrives (prestaged != 0)
    const bool both = JA.exists && JB.exists && cs >= 2;
    const int halfA = cs / 2;
    const bool mineB = both ? (crank >= halfA) : (JB.exists != 0);
    const BoundaryJob J = mineB ? JB : JA;
    const int nside = both ? (mineB ? cs - halfA : halfA) : cs;
    const int csub = both ? (mineB ? crank - halfA : crank) : crank;
    const bool second_pass = JA.exists && JB.exists && cs < 2;   // (a one-CTA cluster serves both boundaries in turn)
    double cmax = -1.0;
    for (int pass = 0; pass < (second_pass ? 2 : 1); ++pass) {
        const BoundaryJob Jp = second_pass ? (pass == 0 ? JA : JB) : J;
        if (!Jp.work) continue;                                    // (uniform over the CTA)
        const int side = Jp.side, c = Jp.c, rc1 = Jp.rc1, rc = Jp.rc;
        const bool corner = Jp.corner != 0;
        const int nc = P.n[c];
        double* XF = C.stg; double* WF = XF + P.d; double* F = WF + P.d;
        const int RE = max(P.Rmax, 32);
        double* T = C.ext; double* DI = T + (i64)RE * RE;
        double* argc = P.arg + P.coreOff[c];
        const bool hasw = (P.kind == KIND_ISING);
        const int nwoff = P.n[1];
        // the mode indices of this CTA: warp w owns x = first(w) .. first(w) + cnt(w) - 1 (contiguous, balanced over all warps)
        const int W = nside * nw, base = nc / W, rem = nc % W;
        const int gw0 = csub * nw;                                 // first global warp of this CTA
        const int cta_first = gw0 * base + min(gw0, rem);
        const int cta_cnt = nw * base + max(0, min(rem - gw0, nw));
        __syncthreads();                                           // ext / stg are free (previous users done)
        if (corner) {
            const int* Lt = P.Lidx + P.offL[c - 1]; const int* Rt = P.Ridx + P.offR[c];
            const i64 oL = P.offL[c - 1] / P.Rmax * P.RT, oR = P.offR[c] / P.Rmax * P.RT;
            for (int pos = threadIdx.x; pos < P.d - 1; pos += blockDim.x) {
                if (P.XLg) {                          // value tables: one round trip instead of index -> value
                    const i64 o = (pos < c - 1) ? oL + (i64)pos * P.RT + (rc1 - 1) : oR + (i64)(pos - (c - 1)) * P.RT + (rc - 1);
                    XF[pos] = __ldcg((pos < c - 1 ? P.XLg : P.XRg) + o); WF[pos] = hasw ? __ldcg((pos < c - 1 ? P.WLg : P.WRg) + o) : 0.0;
                } else {
                    const int idx = (pos < c - 1) ? __ldcg(Lt + (i64)pos * P.Rmax + (rc1 - 1)) : __ldcg(Rt + (i64)(pos - (c - 1)) * P.Rmax + (rc - 1));
                    XF[pos] = P.par[idx - 1]; WF[pos] = hasw ? P.par[nwoff + idx - 1] : 0.0;
                }
            }
        }
        const int rr_ = side == 0 ? rc1 : rc;                      // rank of the recurrence
        const bool flat = rr_ <= 32;                               // thread-per-mode-index solves out of the packed block (ext holds >= 32 x 32 + 32)
        if (!prestaged) {
            if (flat) {
                const double* g = P.inv + (i64)(side == 0 ? c - 1 : c) * P.Rmax * P.Rmax;
                for (int x = threadIdx.x; x < rr_ * rr_; x += blockDim.x) T[x] = LDF(g + x);
                for (int x = threadIdx.x; x < 32 * 32 - rr_ * rr_; x += blockDim.x) T[rr_ * rr_ + x] = 0.0;     // rows >= r of the unrolled solve
                if (side == 1) for (int cc = threadIdx.x; cc < 32; cc += blockDim.x) DI[cc] = cc < rr_ ? 1.0 / LDF(g + (i64)(cc + 1) * (cc + 1) - 1) : 0.0;
            } else if (side == 0) stage_luar_cg(P.inv + (i64)(c - 1) * P.Rmax * P.Rmax, rc1, T);
            else stage_lual_cg(P.inv + (i64)c * P.Rmax * P.Rmax, rc, T, DI);
        }
        if (stage_only) continue;
        __syncthreads();
        double* ST4 = F + P.nmax;                                   // prefix / suffix state of the fixed part of the corner fiber
        if (KIND == KIND_ISINGC && corner) {
            if (threadIdx.x == 0) ising_c_fixed(XF, c - 1, P.d - c, ST4);
            __syncthreads();
        }
        tl_mark(P, 75);
        if (flat) {
            // one thread per mode index: corner value, solve, store
            for (int t = threadIdx.x; t < cta_cnt; t += blockDim.x) {
                const int x = cta_first + t;
                double f = 0.0;
                if (corner) {
                    StagedVals sv;
                    sv.XL = XF; sv.WL = WF; sv.nl = c - 1; sv.rl = 1; sv.i = 1;
                    sv.xj = P.par[x]; sv.wj = hasw ? P.par[nwoff + x] : 0.0;
                    sv.hask = 0; sv.xk = 0.0; sv.wk = 0.0;
                    sv.XR = XF + (c - 1); sv.WR = WF + (c - 1); sv.rr = 1; sv.q = 1;
                    if (KIND == KIND_ISINGC) f = ising_c_eval1(XF, WF, c - 1, P.d - c, ST4, sv.xj, sv.wj);
                    else f = eval_point_wide<KIND>(P, sv, C.A);
                    argc[(rc1 - 1) + (i64)P.Rmax * (x + (i64)nc * (rc - 1))] = f;     // (both members store the same value)
                    cmax = fmax(cmax, fabs(f));
                }
                const double* src = side == 0 ? argc + (i64)P.Rmax * (x + (i64)nc * (rc - 1)) : argc + (rc1 - 1) + (i64)P.Rmax * x;
                const i64 ss = side == 0 ? 1 : (i64)P.Rmax * nc;
                double* dst = side == 0 ? P.rowT + P.coreOff[c] + (i64)nc * (rc - 1) + x : P.col + P.coreOff[c] + (rc1 - 1) + (i64)P.Rmax * x;
                const i64 ds = side == 0 ? (i64)nc * P.Rmax : (i64)P.Rmax * nc;
                if (rr_ <= 16) thread_boundary<16>(side, rr_, src, ss, dst, ds, corner, f, T, DI);
                else thread_boundary<32>(side, rr_, src, ss, dst, ds, corner, f, T, DI);
            }
            tl_mark(P, 76);
            continue;
        }
        if (corner) {
            for (int t = threadIdx.x; t < cta_cnt; t += blockDim.x) {
                const int x = cta_first + t;
                StagedVals sv;
                sv.XL = XF; sv.WL = WF; sv.nl = c - 1; sv.rl = 1; sv.i = 1;
                sv.xj = P.par[x]; sv.wj = hasw ? P.par[nwoff + x] : 0.0;
                sv.hask = 0; sv.xk = 0.0; sv.wk = 0.0;
                sv.XR = XF + (c - 1); sv.WR = WF + (c - 1); sv.rr = 1; sv.q = 1;
                const double f = eval_point_wide<KIND>(P, sv, C.A);
                argc[(rc1 - 1) + (i64)P.Rmax * (x + (i64)nc * (rc - 1))] = f;     // (both members store the same value)
                F[t] = f;
                cmax = fmax(cmax, fabs(f));
            }
            __syncthreads();
        }
        tl_mark(P, 76);
        const int gw = gw0 + wid;
        const int wfirst = gw * base + min(gw, rem), wcnt = base + (gw < rem ? 1 : 0);
        {
            for (int k = 0; k < wcnt; ++k) {
                const int x = wfirst + k;
                const double f = corner ? F[x - cta_first] : 0.0;
                double y[MAXRPL];
                if (side == 0) {
                    const double* src = argc + (i64)P.Rmax * (x + (i64)nc * (rc - 1));
                    double* dst = P.rowT + P.coreOff[c] + (i64)nc * (rc - 1) + x;
                    const i64 de = (i64)nc * P.Rmax;
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) {
                        const int sidx = lane + 32 * u;
                        y[u] = (sidx < rc1) ? ((corner && sidx == rc1 - 1) ? f : __ldcg(src + sidx)) : 0.0;
                    }
                    warp_luar(y, rc1, GSm{T, rc1});
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { const int sidx = lane + 32 * u; if (sidx < rc1) dst[sidx * de] = y[u]; }
                } else {
                    const i64 se = (i64)P.Rmax * nc;
                    const double* src = argc + (rc1 - 1) + (i64)P.Rmax * x;
                    double* dst = P.col + P.coreOff[c] + (rc1 - 1) + (i64)P.Rmax * x;
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) {
                        const int cc = lane + 32 * u;
                        y[u] = (cc < rc) ? ((corner && cc == rc - 1) ? f : __ldcg(src + cc * se)) : 0.0;
                    }
                    warp_lual(y, rc, GSm{T, rc}, DSm{DI});
#pragma unroll
                    for (int u = 0; u < MAXRPL; ++u) { const int cc = lane + 32 * u; if (cc < rc) dst[cc * se] = y[u]; }
                }
            }
        }
    }
    tl_mark(P, 77);
    return cmax;
}

// ----------------------------------------------------------------------------
// grid (CS, nv), cluster (CS, 1, 1), cooperative.  Dynamic shared memory as k_visits.
// ----------------------------------------------------------------------------
#ifndef TTC_SWEEP_MINB
#define TTC_SWEEP_MINB 1
#endif
#ifndef TTC_SWEEP_MAXT
#define TTC_SWEEP_MAXT VISIT_MAXTHREADS
#endif
constexpr int SWEEP_MAXTHREADS = TTC_SWEEP_MAXT;
// MVN: the register-resident evaluation needs ~130 registers for the differences; 192 threads at two CTAs per SM (168 registers)
// evaluates as fast as 255 registers do (measured, config E: 36.8 vs 36.4 ms at 4 x 128) and puts 50 % more threads on a partition
constexpr int SWEEP_MAXTHREADS_MVN = 192;
template <int KIND>
__global__ void __launch_bounds__(KIND == KIND_MVN ? SWEEP_MAXTHREADS_MVN : SWEEP_MAXTHREADS, KIND == KIND_MVN ? 2 : TTC_SWEEP_MINB)
k_sweeps(DevPlan P, int it_last, int maxrank, double small_element, double small_pivot) {
    tl_stamp(P, 36);
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ __align__(16) double smem[];
    __shared__ VisitShared sh;
    __shared__ SweepLocal sl;
    __shared__ __align__(8) unsigned long long tma_bar;
    const int crank = (int)cl.block_rank();
    VisitCtx C;
    C.sh = &sh;
    C.v = P.v0 + blockIdx.y;
    C.lo = P.own[C.v]; C.hi = P.own[C.v + 1];
    // MVN: the d x d matrix in shared memory whenever the host made room for it (auxsm_p; the per-sweep kernels keep it in
    // global memory beyond 8 KB because they run 4 CTAs per SM)
    const int auxd = KIND == KIND_MVN ? max(P.auxsm, P.auxsm_p) : 0;
    if (KIND == KIND_MVN && auxd >= P.d * P.d) {
        for (int x = threadIdx.x; x < P.d * P.d; x += blockDim.x) smem[x] = P.aux[P.d + x];
        __syncthreads();
        C.A = smem;
    } else C.A = stage_aux<KIND>(P, smem);
    C.xs = smem + auxd;
    C.pre = C.xs + P.Rmax;
    C.ext = C.pre + 4 * (i64)P.Rmax;
    C.stg = C.ext + (i64)max(P.Rmax, 32) * max(P.Rmax, 32) + max(P.Rmax, 32);
    C.stg = (double*)(((unsigned long long)C.stg + 15ULL) & ~15ULL);          // TMA destinations are 16-byte aligned
    C.ibuf = (int*)(C.stg + P.stage_max);
    C.phase = 0;
    C.bar = &tma_bar; C.bar_parity = 0;
    const int v = C.v, lo = C.lo, hi = C.hi;
    const bool multi = P.nproc > 1;
    SweepMail* mail = P.mail;
    const unsigned long long seq0 = P.ctrl->run_serial * 65536ULL;
    if (threadIdx.x == 0) {
        sh.S = P.st[v];
        sl.rkL = P.rk[lo - 1]; sl.rkR = P.rk[hi];
        sl.strike = 0; sl.ready = 0;
        mbar_init(&tma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // value tables of the bonds this partition reads (lo-1 .. hi) from the index tables of the initial cross; neighbours
    // write the same values into the bonds they share
    {
        const bool hasw = (P.kind == KIND_ISING);
        const int nwoff = P.n[1];
        const int gtid = crank * blockDim.x + threadIdx.x, gthreads = (int)cl.num_blocks() * blockDim.x;
        for (int b = lo - 1; b <= hi; ++b) {
            const int rb = P.rk[b];
            const int* Lt = P.Lidx + P.offL[b]; const int* Rt = P.Ridx + P.offR[b];
            const i64 oL = P.offL[b] / P.Rmax * P.RT, oR = P.offR[b] / P.Rmax * P.RT;
            for (int x = gtid; x < b * rb; x += gthreads) {
                const int pos = x / rb, t = x - pos * rb, idx = Lt[(i64)pos * P.Rmax + t];
                P.XLg[oL + (i64)pos * P.RT + t] = P.par[idx - 1]; if (hasw) P.WLg[oL + (i64)pos * P.RT + t] = P.par[nwoff + idx - 1];
            }
            for (int x = gtid; x < (P.d - b) * rb; x += gthreads) {
                const int pos = x / rb, t = x - pos * rb, idx = Rt[(i64)pos * P.Rmax + t];
                P.XRg[oR + (i64)pos * P.RT + t] = P.par[idx - 1]; if (hasw) P.WRg[oR + (i64)pos * P.RT + t] = P.par[nwoff + idx - 1];
            }
        }
        fence_proxy_async();
        __threadfence();
    }
    cl.sync();
    if (threadIdx.x == 0) sl.pending = 0;
    // Close of sweep `cit` (MAX allreduce dmrgg.f90:852-870, sweep record :961-1008, exit test :1010-1019), identical in every
    // CTA of every cluster.  It is taken LATE: inside the next sweep's first visit, right before its accept test (the visit's
    // lottery and fibers touch nothing outside the cluster), so the wait for everybody's records hides behind them; `spec` says
    // that such a visit is under way: if the exit test fires, the partition state goes back to the end of sweep `cit`.
    auto close_sweep = [&](int cit, bool spec) {
        for (int u = threadIdx.x; u < P.P; u += blockDim.x) {
        sweep_wait(P, &mail[u].flag2, seq0 + (unsigned long long)cit, multi);
        const volatile SweepRec& R = mail[u].rec[cit & 1];
        sl.r_amax1[u & 63] = R.amax1; sl.r_pmax[u & 63] = R.pivotmax; sl.r_pmin[u & 63] = R.pivotmin;
        sl.r_amax2[u & 63] = R.amax2; sl.r_neval[u & 63] = R.neval; sl.r_err[u & 63] = R.error;
    }
    __syncthreads();
    tl_mark(P, 73);
    // ---- MAX allreduce (dmrgg.f90:852-870), sweep record (:961-1008), exit test (:1010-1019); identical in every CTA
    if (threadIdx.x == 0) {
        double c1 = sl.r_amax1[0], c2 = sl.r_pmax[0], c3 = (sl.r_pmin[0] > 0.0) ? -sl.r_pmin[0] : -999e9;
        long long ne = sl.r_neval[0];
        int err = sl.r_err[0];
        for (int u = 1; u < P.P; ++u) {
            c1 = fmax(c1, sl.r_amax1[u]); c2 = fmax(c2, sl.r_pmax[u]);
            c3 = fmax(c3, (sl.r_pmin[u] > 0.0) ? -sl.r_pmin[u] : -999e9);
            ne += sl.r_neval[u]; err |= sl.r_err[u];
        }
        const double s_pmax = c2, s_pmin = (-c3 == 999e9) ? -1.0 : -c3;
        const double s_amax = fmax(c1, sl.r_amax2[0]);     // st[0].amax after the exchange
        int ready = 0;
        if (maxrank > 0) ready = (cit + 1 >= maxrank);
        if (P.ctrl->has_accuracy) {
            if (s_pmax <= P.ctrl->accuracy * s_amax) sl.strike += 1; else sl.strike = 0;
            ready = ready || (sl.strike >= 3);
        }
        if (err) ready = 1;
        sl.ready = ready;
        if (ready && spec) sh.S = sl.saved;          // the visit under way is abandoned
        VState& St = sh.S;
        St.amax = fmax(c1, St.amax);                 // the allreduced value, then this partition's corner fibers (and, speculatively, the next visit's)
        St.pivotmax_prev = s_pmax;                   // dmrgg.f90:961
        St.pivotmax = -1.0; St.pivotmin = -1.0;      // dmrgg.f90:326-327 of the next sweep
        if (crank == 0) {
            for (int x = lo; x <= hi - 1; ++x) { const int r = LDF(P.rk + x); P.rklog[(i64)cit * (P.d + 1) + x] = r; P.rks[x] = r; }
            if (v == 0) P.rklog[(i64)cit * (P.d + 1)] = 1;
            if (v == P.P - 1) P.rklog[(i64)cit * (P.d + 1) + P.d] = 1;
            if (v == P.v0) {                         // one record per process
                SweepOut& O = P.slog[cit];
                O.neval = ne; O.amax = s_amax; O.pivotmax = s_pmax; O.pivotmin = s_pmin;
                O.t_ns = globaltimer_ns() - P.ctrl->t0_ns; O.valid = 1; O.pad = 0;
                if (err && !P.ctrl->error) P.ctrl->error = err;
            }
        }
    }
        __syncthreads();
        tl_mark(P, 74);
    };
    int it = 1;
    for (; it <= it_last; ++it) {
        const int dir = 2 - (it & 1);                  // dmrgg.f90:317
        const int par = it & 1;
        const unsigned long long seq = seq0 + (unsigned long long)it;
        C.rkL = sl.rkL; C.rkR = sl.rkR;
        auto hook = [&]() -> bool {
            if (sl.pending == 0) return false;           // (first sweep: nothing to close)
            close_sweep(it - 1, true);
            if (threadIdx.x == 0) sl.pending = 0;
            return sl.ready != 0;
        };
        if (visit_list<KIND, true, true>(P, cl, C, it, dir, small_element, small_pivot, hook)) { it -= 1; break; }      // ends with a cluster barrier

        // ---- several processes: the first / last partition of a process pushes what its foreign neighbour needs straight into
        // that process's window over NVLink (the tape of dmrgg.f90:763-850 reduced to the one table column the neighbour will
        // dereference, and the blocks of :872-958 / dmrggmp.f90:572-629), then fences at system scope
        const bool edgeL = multi && v == P.v0 && P.prank > 0, edgeR = multi && v == P.v0 + P.nv - 1 && P.prank < P.nproc - 1;
        if (edgeL || edgeR) {
            const int gtid = crank * blockDim.x + threadIdx.x, gthreads = (int)cl.num_blocks() * blockDim.x;
            if (edgeL && C.upd_first) {               // to the LEFT process: new slice of core lo, new column of the R table of bond lo
                const int c = lo, n = P.n[c], rc = LDF(P.rk + c);
                char* dstw = P.peer_win[P.prank - 1] + P.win_sr + (long long)par * P.slab_r_bytes;
                double* dst = (double*)dstw;
                const double* slab = P.arg + P.coreOff[c] + (i64)P.Rmax * n * (rc - 1);
                for (int x = gtid; x < P.Rmax * n; x += gthreads) dst[x] = LDF(slab + x);
                int* dcol = (int*)(dst + (i64)P.Rmax * P.nmax);
                const int* Rt = P.Ridx + P.offR[c];
                for (int pos = gtid; pos < P.d - c; pos += gthreads) dcol[pos] = LDF(Rt + (i64)pos * P.Rmax + (rc - 1));
            }
            if (edgeR && C.upd_last) {                // to the RIGHT process: new row of core hi, L-table column and packed-LU row of bond hi-1
                const int c = hi, n = P.n[c], rc1 = LDF(P.rk + c - 1), rq = sl.rkR;
                char* dstw = P.peer_win[P.prank + 1] + P.win_sl + (long long)par * P.slab_l_bytes;
                double* dst = (double*)dstw;
                const double* a = P.arg + P.coreOff[c] + (rc1 - 1);
                for (int x = gtid; x < n * rq; x += gthreads) dst[x] = LDF(a + (i64)P.Rmax * x);
                int* dcol = (int*)(dst + (i64)P.nmax * P.Rmax);
                const int* Lt = P.Lidx + P.offL[c - 1];
                for (int pos = gtid; pos < c - 1; pos += gthreads) dcol[pos] = LDF(Lt + (i64)pos * P.Rmax + (rc1 - 1));
                double* dlu = dst + (i64)P.nmax * P.Rmax + P.d;
                const double* g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax + (i64)(rc1 - 1) * (rc1 - 1);
                for (int x = gtid; x < 2 * (rc1 - 1) + 1; x += gthreads) dlu[x] = LDF(g + x);
            }
            __threadfence_system();
            cl.sync();
        }
        // ---- announce the bond visits: everything this cluster wrote is ordered before the flag (the cluster barrier
        // synchronises the writers with this thread, the release is cumulative)
        if (crank == 0 && threadIdx.x < 3) {
            // thread 0: this process's window; 1 / 2: the left / right process's window when the neighbour partition lives there
            const int g = threadIdx.x == 0 ? P.prank : (threadIdx.x == 1 ? P.prank - 1 : P.prank + 1);
            const bool go = threadIdx.x == 0 || (threadIdx.x == 1 ? edgeL : edgeR);
            if (go) {
                SweepMail* mb = (multi ? peer_ptr<SweepMail>(P, g, P.win_mail) : mail) + v;
                mb->upd1[par][0] = C.upd_first; mb->upd1[par][1] = C.upd_last;
                if (multi) { __threadfence_system(); st_release<true>(&mb->flag1, seq); }
                else { __threadfence(); st_release<false>(&mb->flag1, seq); }
            }
        }
        if (threadIdx.x == 0) sl.amax1 = sh.S.amax;
        tl_mark(P, 70);
        // ---- while the flags travel: stage this CTA's packed LU for the exchange (own data: inv(lo) resp. inv(hi-1))
        BoundaryJob JA = {v > 0, v > 0, 1, lo, 0, 0, 0}, JB = {v < P.P - 1, v < P.P - 1, 0, hi, 0, 0, 0};     // (work = exists while staging)
        if (v > 0) JA.rc = LDF(P.rk + lo);
        if (v < P.P - 1) JB.rc1 = LDF(P.rk + hi - 1);
        const int prestaged = (int)cl.num_blocks() >= 2 ? 1 : 0;
        if (prestaged) exchange_boundaries<KIND>(P, cl, C, JA, JB, 1, 0);
        // ---- wait for the two neighbours (each CTA polls for itself: no extra cluster barrier)
        if (threadIdx.x < 2) {
            const int u = threadIdx.x == 0 ? v - 1 : v + 1;
            int upd = 0;
            if (u >= 0 && u < P.P) {
                sweep_wait(P, &mail[u].flag1, seq, multi);
                upd = *(volatile int*)&mail[u].upd1[par][threadIdx.x == 0 ? 1 : 0];
            }
            if (threadIdx.x == 0) sl.updL = upd; else sl.updR = upd;
        }
        __syncthreads();
        tl_mark(P, 71);
        const int updL = sl.updL, updR = sl.updR;
        if ((edgeL && updL) || (edgeR && updR)) {       // drop what the foreign neighbour pushed into place (uniform over the cluster)
            const int gtid = crank * blockDim.x + threadIdx.x, gthreads = (int)cl.num_blocks() * blockDim.x;
            if (edgeL && updL) {                        // from the LEFT process: new row of core lo | L column of bond lo-1 | packed-LU row of inv(lo-1)
                const int c = lo, n = P.n[c], t = sl.rkL, rq = LDF(P.rk + c) - C.upd_first;      // the sender saw bond c at its sweep-start rank
                const double* src = (const double*)((const char*)P.mail + (P.win_sl - P.win_mail) + (long long)par * P.slab_l_bytes);
                double* a = P.arg + P.coreOff[c] + t;
                for (int x = gtid; x < n * rq; x += gthreads) a[(i64)P.Rmax * x] = __ldcg(src + x);
                const int* scol = (const int*)(src + (i64)P.nmax * P.Rmax);
                int* Lt = P.Lidx + P.offL[c - 1];
                const bool hasw = (P.kind == KIND_ISING); const int nwoff = P.n[1];
                const i64 oL = P.offL[c - 1] / P.Rmax * P.RT;
                for (int pos = gtid; pos < c - 1; pos += gthreads) {
                    const int idx = __ldcg(scol + pos);
                    Lt[(i64)pos * P.Rmax + t] = idx;
                    P.XLg[oL + (i64)pos * P.RT + t] = P.par[idx - 1]; if (hasw) P.WLg[oL + (i64)pos * P.RT + t] = P.par[nwoff + idx - 1];
                }
                const double* slu = src + (i64)P.nmax * P.Rmax + P.d;
                double* g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax + (i64)t * t;
                for (int x = gtid; x < 2 * t + 1; x += gthreads) g[x] = __ldcg(slu + x);
            }
            if (edgeR && updR) {                        // from the RIGHT process: new slice of core hi | R column of bond hi
                const int c = hi, n = P.n[c], t = sl.rkR, ri = LDF(P.rk + c - 1) - C.upd_last;   // the sender saw bond c-1 at its sweep-start rank
                const double* src = (const double*)((const char*)P.mail + (P.win_sr - P.win_mail) + (long long)par * P.slab_r_bytes);
                double* slab = P.arg + P.coreOff[c] + (i64)P.Rmax * n * t;
                for (int x = gtid; x < P.Rmax * n; x += gthreads) { if (x % P.Rmax < ri) slab[x] = __ldcg(src + x); }
                const int* scol = (const int*)(src + (i64)P.Rmax * P.nmax);
                int* Rt = P.Ridx + P.offR[c];
                const bool hasw = (P.kind == KIND_ISING); const int nwoff = P.n[1];
                const i64 oR = P.offR[c] / P.Rmax * P.RT;
                for (int pos = gtid; pos < P.d - c; pos += gthreads) {
                    const int idx = __ldcg(scol + pos);
                    Rt[(i64)pos * P.Rmax + t] = idx;
                    P.XRg[oR + (i64)pos * P.RT + t] = P.par[idx - 1]; if (hasw) P.WRg[oR + (i64)pos * P.RT + t] = P.par[nwoff + idx - 1];
                }
            }
            fence_proxy_async();
            cl.sync();
        }
        // ---- this partition's half of the exchange on its two boundaries
        int ncorner = 0;
        if (v > 0) {                                    // RIGHT member of boundary v-1: core lo
            JA.work = updL; JA.rc1 = sl.rkL + updL; JA.corner = updL && C.upd_first;
            if (JA.corner) ncorner += P.n[lo];
        }
        if (v < P.P - 1) {                              // LEFT member of boundary v: core hi
            JB.work = updR; JB.rc = sl.rkR + updR; JB.corner = C.upd_last && updR;
            if (JB.corner) ncorner += P.n[hi];
        }
        double cmax = exchange_boundaries<KIND>(P, cl, C, JA, JB, 0, prestaged);
        if (ncorner) {                                  // uniform over the cluster
            Partial dummy = amax_init();
            cluster_fold(cl, sh, C.phase, cmax, dummy);
            if (threadIdx.x == 0) { sh.S.amax = fmax(sh.S.amax, cmax); sh.S.neval += ncorner; }
        } else {
            cl.sync();                                  // the appended factor entries are read by other CTAs in the next sweep
        }
        tl_mark(P, 72);
        // ---- publish the scalars, wait for everybody's
        if (crank == 0 && threadIdx.x < P.nproc) {     // one thread per destination window
            SweepMail* mb = (multi ? peer_ptr<SweepMail>(P, threadIdx.x, P.win_mail) : mail) + v;
            SweepRec& R = mb->rec[par];
            R.amax1 = sl.amax1; R.pivotmax = sh.S.pivotmax; R.pivotmin = sh.S.pivotmin;
            R.amax2 = sh.S.amax; R.neval = sh.S.neval; R.error = *(volatile int*)&P.ctrl->error; R.pad = 0;
            if (multi) { __threadfence_system(); st_release<true>(&mb->flag2, seq); }
            else { __threadfence(); st_release<false>(&mb->flag2, seq); }
        }
        if (threadIdx.x == 0) { sl.rkL += updL; sl.rkR += updR; sl.pending = it; sl.saved = sh.S; }
        __syncthreads();
    }
    if (it > it_last) { it = it_last; if (sl.pending) close_sweep(it, false); }
    if (crank == 0 && threadIdx.x == 0) {
        P.st[v] = sh.S;
        if (v == P.v0) { P.ctrl->nsweeps = it; P.ctrl->it = it + 1; P.ctrl->strike = sl.strike; P.ctrl->ready = 1; }
    }
    if (multi) {
        // every process ends with the complete pivot tape and rank log (the reference's tape reaches every rank, dmrgg.f90:763-850):
        // each partition pushes its records of all sweeps into every other window, then a last flag round
        if (crank == 0) {
            for (int g = 0; g < P.nproc; ++g) {
                if (g == P.prank) continue;
                VisitOut* dv = peer_ptr<VisitOut>(P, g, P.win_vlog);
                int* dr = peer_ptr<int>(P, g, P.win_rklog);
                const int nrec = it * P.maxnb;
                for (int x = threadIdx.x; x < nrec * (int)(sizeof(VisitOut) / 8); x += blockDim.x) {
                    const int rec = x / (int)(sizeof(VisitOut) / 8), w = x - rec * (int)(sizeof(VisitOut) / 8);
                    const i64 idx = (i64)rec * P.P + v;
                    ((unsigned long long*)(dv + idx))[w] = ((const unsigned long long*)(P.vlog + idx))[w];
                }
                const int nbo = hi - lo + (v == 0 ? 1 : 0) + (v == P.P - 1 ? 1 : 0), b0 = (v == 0) ? 0 : lo;
                for (int x = threadIdx.x; x < it * nbo; x += blockDim.x) {
                    const int sw = x / nbo + 1, bx = b0 + (x - (sw - 1) * nbo);
                    dr[(i64)sw * (P.d + 1) + bx] = P.rklog[(i64)sw * (P.d + 1) + bx];
                }
            }
            __threadfence_system();
            __syncthreads();
            const unsigned long long seqf = seq0 + 65535ULL;
            if (threadIdx.x < P.nproc) st_release<true>(&peer_ptr<SweepMail>(P, threadIdx.x, P.win_mail)[v].flag1, seqf);
            for (int u = threadIdx.x; u < P.P; u += blockDim.x) sweep_wait(P, &mail[u].flag1, seqf, true);
            __syncthreads();
            if (v == P.v0) for (int x = threadIdx.x; x <= P.d; x += blockDim.x) { const int r = __ldcg(P.rklog + (i64)it * (P.d + 1) + x); P.rk[x] = r; P.rks[x] = r; }
        }
    }
}

// ----------------------------------------------------------------------------
// Per-sweep quadrature values after the loop.  The contracted cores ttqq have been formed and dtt_lua'd at the FINAL
// ranks (k_quad_contract_sm, k_quad_lua_all); sweep s uses their leading rklog[s] blocks.
// k_quad_chain_all: grid (nv, S): chain product of partition v0 + x at sweep y + 1 -> chainS[(sweep - 1) * (P + 1) + v]
// k_quad_tree_all : grid (S): the reference's binary tree over partitions (dmrgg.f90:1355-1405) -> slog[sweep].val
// dynamic smem: 3 Rmax^2 doubles.
// ----------------------------------------------------------------------------
__global__ void k_quad_lua_all(DevPlan P) {
    extern __shared__ double smem[];
    const int p = P.c_lo + blockIdx.x;
    const int r0 = P.rk[p - 1], r1 = P.rk[p];
    double* M = smem; double* TL = M + r0 * r1; double* TR = TL + r0 * r0; double* DI = TR + r1 * r1;
    double* gm = P.ttqq + (i64)p * P.Rmax * P.Rmax;
    for (int x = threadIdx.x; x < r0 * r1; x += blockDim.x) { int k = x / r0, i = x - k * r0; M[x] = gm[i + (i64)P.Rmax * k]; }
    stage_luar(P.inv + (i64)(p - 1) * P.Rmax * P.Rmax, r0, TL);
    if (p < P.d) stage_lual(P.inv + (i64)p * P.Rmax * P.Rmax, r1, TR, DI);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int k = wid; k < r1; k += nw) {
        double y[MAXRPL];
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; y[t] = (sidx < r0) ? M[sidx + r0 * k] : 0.0; }
        warp_luar(y, r0, GSm{TL, r0});
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; if (sidx < r0) M[sidx + r0 * k] = y[t]; }
    }
    __syncthreads();
    if (p < P.d) {
        for (int i = wid; i < r0; i += nw) {
            double y[MAXRPL];
#pragma unroll
            for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; y[t] = (c < r1) ? M[i + r0 * c] : 0.0; }
            warp_lual(y, r1, GSm{TR, r1}, DSm{DI});
#pragma unroll
            for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; if (c < r1) M[i + r0 * c] = y[t]; }
        }
        __syncthreads();
    }
    for (int x = threadIdx.x; x < r0 * r1; x += blockDim.x) { int k = x / r0, i = x - k * r0; gm[i + (i64)P.Rmax * k] = M[x]; }
}
// epilogue of a kernel whose results other processes wait for (every thread of every CTA calls it): once the whole grid's
// peer stores are fenced, the last CTA stores this phase's sequence number into every process's flag slot (cf. mp_publish)
__device__ __forceinline__ void mp_publish_grid(const DevPlan& P, int phase) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(P.tickets + P.P, 1u);
        s_last = (t == gridDim.x * gridDim.y * gridDim.z - 1);
        if (s_last) P.tickets[P.P] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    const unsigned long long seq = mp_seq(P, phase);
    for (int q = threadIdx.x; q < P.nproc; q += blockDim.x)
        st_release_sys((unsigned long long*)(P.peer_win[q] + P.win_flg) + (long long)(phase - 1) * P.nproc + P.prank, seq);
}
__global__ void k_quad_chain_all(DevPlan P, double* chainS) {
    extern __shared__ double smem[];
    const int v = P.v0 + blockIdx.x, sw = blockIdx.y + 1;
    if (sw <= P.ctrl->nsweeps) {
        const int first = P.own[v];
        int last = P.own[v + 1] - 1;
        if (v == P.P - 1) last = P.d;
        const int ld = P.Rmax;
        const int* rq = P.rklog + (i64)sw * (P.d + 1);
        const i64 msz = (i64)ld * ld;
        double* cur = smem; double* nxt = smem + msz; double* B = smem + 2 * msz;
        const int m = rq[first - 1];
        mat_load_sm(P.ttqq + (i64)first * msz, m, rq[first], ld, cur, ld);
        __syncthreads();
        for (int p = first + 1; p <= last; ++p) {
            mat_load_sm(P.ttqq + (i64)p * msz, rq[p - 1], rq[p], ld, B, ld);
            __syncthreads();
            mat_mul_sm(cur, m, rq[p - 1], B, rq[p], nxt, ld);
            __syncthreads();
            double* t = cur; cur = nxt; nxt = t;
        }
        const i64 off = ((i64)(sw - 1) * (P.P + 1) + v) * msz;
        const int nl = rq[last];
        for (int g = 0; g < P.nproc; ++g) {          // the chain product goes to every process (each runs the tree, dmrgg.f90:1355-1405)
            double* out = (P.nproc > 1 ? (double*)(P.peer_win[g] + P.win_chs) : chainS) + off;
            for (int e = threadIdx.x; e < m * nl; e += blockDim.x) { int j = e / m, i = e - j * m; out[i + (i64)ld * j] = cur[i + ld * j]; }
        }
        if (P.P == 1 && threadIdx.x == 0) P.slog[sw].val = cur[0];
    }
    if (P.nproc > 1) mp_publish_grid(P, 2);
}
__global__ void k_quad_tree_all(DevPlan P, double* chainS) {
    extern __shared__ double smem[];
    const int sw = blockIdx.x + 1;
    if (P.nproc > 1) mp_wait(P, 2);                  // every process's chain products have landed in this window
    if (sw > P.ctrl->nsweeps || P.P == 1) return;
    const int ld = P.Rmax;
    const int* rq = P.rklog + (i64)sw * (P.d + 1);
    const i64 msz = (i64)ld * ld;
    double* ch = chainS + (i64)(sw - 1) * (P.P + 1) * msz;
    double* A = smem; double* B = smem + msz; double* Cm = smem + 2 * msz;
    for (int q = 1; q < P.P; q *= 2) {
        for (int me = 0; me + q < P.P; me += 2 * q) {
            const int her = me + q;
            int herend = her + q; if (herend > P.P) herend = P.P;
            const int m = rq[P.own[me] - 1], kd = rq[P.own[her] - 1];
            const int n = (herend == P.P) ? rq[P.d] : rq[P.own[herend] - 1];
            mat_load_sm(ch + (i64)me * msz, m, kd, ld, A, ld);
            mat_load_sm(ch + (i64)her * msz, kd, n, ld, B, ld);
            __syncthreads();
            mat_mul_sm(A, m, kd, B, n, Cm, ld);
            __syncthreads();
            double* out = ch + (i64)me * msz;
            for (int e = threadIdx.x; e < m * n; e += blockDim.x) { int j = e / m, i = e - j * m; out[i + (i64)ld * j] = Cm[i + ld * j]; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) P.slog[sw].val = ch[0];
}

// ----------------------------------------------------------------------------
// Device-side initial cross (dmrgg.f90:150-270) for the persistent schedule: no host round trip between the kernels.
// k_init_pick : first-index argmax of |b| over the nn x snum diagonal search points (idamax + MAXLOC, :179-203), the
//               initial multi-index ind0, pivot 1 of every bond (vip, ranks, index tables)                       <<<1, 1024>>>
// k_init_state: after k_init_cross / k_init_factors: per-partition amax / neval (:193-232), the value of the rank-1 train
//               (:250-270) and the '0::' record slog[0]                                                          <<<1, 256>>>
// scal: [0] = gmax.
// ----------------------------------------------------------------------------
__global__ void k_init_pick(DevPlan P, int nn, int snum, const double* b, double* scal, int* ind0 /*[d + 2]*/) {
    __shared__ Partial shp[32];
    __shared__ int s_ind[MAXD_LOCAL * 8 + 2];
    Partial best = amax_init();
    for (int x = threadIdx.x; x < nn * snum; x += blockDim.x) amax_take(best, b[x], x);
    best = amax_block(best, shp);
    __shared__ long long s_x;
    if (threadIdx.x == 0) { s_x = best.idx; scal[0] = best.absv; }
    __syncthreads();
    const int gilot = (int)s_x + 1;
    const int sft = (gilot - 1) / nn, k = (gilot - 1) % nn + 1;
    const int d = P.d, Rmax = P.Rmax;
    for (int p = threadIdx.x; p <= d + 1; p += blockDim.x) {
        const int v = (p >= 1 && p <= d) ? (k - 1 + sft * (p - 1)) % P.n[p] + 1 : 1;
        ind0[p] = v;
        if (p < (int)(sizeof(s_ind) / sizeof(int))) s_ind[p] = v;
    }
    __syncthreads();
    auto I0 = [&](int p) { return p < (int)(sizeof(s_ind) / sizeof(int)) ? s_ind[p] : ((p >= 1 && p <= d) ? (k - 1 + sft * (p - 1)) % P.n[p] + 1 : 1); };
    for (int p = threadIdx.x; p <= d; p += blockDim.x) {
        int* t = P.vip + (i64)p * Rmax * 4;
        t[0] = 1; t[3] = 1;
        t[1] = (p >= 1 && p <= d - 1) ? I0(p) : 1;
        t[2] = (p >= 1 && p <= d - 1) ? I0(p + 1) : 1;
        P.rk[p] = 1; P.rks[p] = 1;
    }
    if (threadIdx.x == 0) { P.rk[d + 1] = 1; P.rks[d + 1] = 1; }
    // pivot 1 of the flat tables: L(p)(pos, 0) = ind0(pos + 1), pos < p;  R(p)(pos, 0) = ind0(p + pos + 1), pos < d - p
    for (int x = threadIdx.x; x < (d + 1) * d; x += blockDim.x) {
        const int p = x / d, pos = x - p * d;
        if (pos < p) P.Lidx[P.offL[p] + (i64)pos * Rmax] = I0(pos + 1);
        if (pos < d - p) P.Ridx[P.offR[p] + (i64)pos * Rmax] = I0(p + pos + 1);
    }
}
__global__ void k_init_state(DevPlan P, int nn, int snum, const double* scal, const int* ind0, int has_quad, int stage_w) {
    extern __shared__ double fibs[];                      // fibs[(p-1) * nmax + j]: the initial fiber of every core, staged by all threads
    __shared__ double s_dot[MAXD_LOCAL * 8 + 2], s_part[64], s_am[MAXD_LOCAL * 8 + 2];
    const int d = P.d, NP = P.P;
    const double gmax = scal[0];
    double* wsm = fibs + (size_t)d * P.nmax;              // the weights beside them (stage_w): the ordered sums below never wait for HBM
    for (int x = threadIdx.x; x < d * P.nmax; x += blockDim.x) {
        const int p = x / P.nmax + 1, j = x - (p - 1) * P.nmax;
        fibs[x] = (j < P.n[p]) ? P.arg[P.coreOff[p] + (i64)P.Rmax * j] : 0.0;
        if (has_quad && stage_w) wsm[x] = (j < P.n[p]) ? P.quadw[P.quadOff[p] + j] : 0.0;
    }
    __syncthreads();
    // ddot(fiber(p), quad(p)) per core, sequential in j like the reference's ddot (out of shared memory: no load latency on the chain)
    for (int p = 1 + threadIdx.x; p <= d; p += blockDim.x) {
        double t = 0.0;
        if (has_quad && p < (int)(sizeof(s_dot) / sizeof(double))) {
            const double* a = fibs + (size_t)(p - 1) * P.nmax;
            const double* w = stage_w ? wsm + (size_t)(p - 1) * P.nmax : P.quadw + P.quadOff[p];
            for (int j = 0; j < P.n[p]; ++j) t = t + a[j] * w[j];
            s_dot[p] = t;
        }
    }
    // max |fiber(p)| per core: a warp per core (a maximum does not depend on the order)
    for (int p = 1 + (int)(threadIdx.x >> 5); p <= d && p < (int)(sizeof(s_am) / sizeof(double)); p += (int)(blockDim.x >> 5)) {
        const double* a = fibs + (size_t)(p - 1) * P.nmax;
        double am = 0.0;
        for (int j = threadIdx.x & 31; j < P.n[p]; j += 32) am = fmax(am, fabs(a[j]));
        for (int o = 16; o; o >>= 1) am = fmax(am, __shfl_xor_sync(0xffffffffu, am, o));
        if ((threadIdx.x & 31) == 0) s_am[p] = am;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < NP; v += blockDim.x) {
        const int s0 = (int)((double)snum * (double)v / NP), s1 = (v + 1 == NP) ? snum : (int)((double)snum * (double)(v + 1) / NP);
        double am = gmax;
        long long ne = (long long)nn * (s1 - s0);
        for (int p = P.own[v]; p <= P.own[v + 1]; ++p) {
            ne += P.n[p];
            if (p < (int)(sizeof(s_am) / sizeof(double))) am = fmax(am, s_am[p]);
            else { const double* a = fibs + (size_t)(p - 1) * P.nmax; for (int j = 0; j < P.n[p]; ++j) am = fmax(am, fabs(a[j])); }
        }
        VState S;
        S.ii = S.jj = S.kk = S.qq = 0; S.pivot = 0.0; S.done = S.havecol = S.haverow = S.crs = 0; S.upd = 0; S.pad0 = 0;
        S.amax = am; S.pivotmax = -1.0; S.pivotmin = -1.0; S.pivotmax_prev = am; S.neval = ne; S.rng_k = 0ULL;
        P.st[v] = S;
        double x = 1.0;
        if (has_quad) {
            for (int p = P.own[v]; p <= P.own[v + 1] - 1; ++p) x = x * s_dot[p] / fibs[(size_t)(p - 1) * P.nmax + (ind0[p] - 1)];
            if (v == NP - 1) x = x * s_dot[d];
        }
        if (v < 64) s_part[v] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double val = 0.0;
        if (has_quad) { val = s_part[0]; for (int v = 1; v < NP; ++v) val = val * s_part[v]; }
        long long ne = 0;
        for (int v = 0; v < NP; ++v) ne += P.st[v].neval;
        SweepOut& O = P.slog[0];
        O.val = val; O.neval = ne; O.amax = P.st[0].amax; O.pivotmax = -1.0; O.pivotmin = -1.0; O.t_ns = 0; O.valid = 1; O.pad = 0;
        for (int x = 0; x <= d; ++x) P.rklog[x] = 1;
    }
}

}  // namespace ttc
