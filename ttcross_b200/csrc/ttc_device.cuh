// =============================================================================
// ttc_device.cuh — device data layout + sm_100a kernels of the TT-cross sweep.
//
// Everything here is compiled with -fmad=false: the reference arithmetic is plain
// IEEE double without contraction, and every reduction below keeps the
// reference's (netlib BLAS) summation order so results are bit-identical to the
// CPU oracle for the +-*/ integrands.
//
// HBM layout (fixed leading dimensions, nothing is ever reallocated; the reference
// reallocates and copies every block on every rank increment, dmrgg.f90:638-753):
//   arg(p), col(p) : element (i,j,k) at (i-1) + Rmax*((j-1) + n(p)*(k-1))        [reference order, padded i]
//   rowT(p)        : element (s,k,q) at (k-1) + n(p)*((q-1) + Rmax*(s-1))          [s slowest: residual/luar
//                    loops over s read coalesced in k]
//   inv(p)         : packed incremental LU, (dmrgg.f90:650-660), Rmax^2 doubles per bond
//   Lidx(p)/Ridx(p): flat left/right multi-indices of every pivot of bond p, position-major, so the
//                    pointer chase of dmrgg_fun (dmrgg.f90:1062-1075) becomes two table rows
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include "../../include/ttc_detexp.h"

namespace ttc {

typedef long long i64;

enum { KIND_ISING = 1, KIND_STDNORM = 4, KIND_MVN = 5, KIND_COSCOEF = 6,
       KIND_ISINGC = 7 };   // compile-time tag of the bond-visit kernels for Ising with id = 1 (P.kind stays KIND_ISING)
constexpr int MAXD_LOCAL = 64;     // integrands gather node values into registers/local memory up to this d
constexpr int GMAX = 1024;         // max CTAs per virtual rank in any reducing kernel

struct Partial { double absv; double val; i64 idx; };

struct VState {            // one per virtual rank, device resident
    int ii, jj, kk, qq;
    double pivot;
    int done, havecol, haverow, crs;
    int upd, pad0;
    double amax, pivotmax, pivotmin, pivotmax_prev;
    i64 neval;
    unsigned long long rng_k;   // position in this virtual rank's uniform stream
};
struct VisitOut {          // what the host reads back after a bond visit
    int active, upd, bond, ii, jj, kk, qq, pad;
    double pivot;
};
struct SweepOut {
    double val, amax, pivotmax, pivotmin;
    i64 neval;
    unsigned long long t_ns;    // %globaltimer at the end of the sweep
    int valid, pad;
};
struct Ctrl {                   // device-side control of the asynchronous sweep loop
    int ready;                  // exit condition reached (dmrgg.f90:1010-1019): later sweeps are no-ops
    int strike;
    int error;                  // 1: rank capacity exceeded
    int nsweeps;                // sweeps actually performed
    unsigned long long t0_ns;
    int it;                     // current sweep (advanced by k_sweep_begin, so captured graphs carry no sweep number)
    unsigned long long run_serial;   // ttc_dmrgg call number on this handle (sequence numbers of the peer-memory exchange)
    unsigned long long quad_serial;  // collective ttc_quad call number
    int has_accuracy;
    double accuracy;
    unsigned long long seed;    // uniform stream seed
    int itq;                    // overlapped quadrature (second stream): sweeps whose quadrature value has been recorded
    int pad_itq;
};
// exact restatement of lottery2's cumulative weights (rnd.f90:115-125) for 0/1 weights, see build_segments()
struct LotSeg { long long M; long long k; int c0; int J; int E; int pad; };
constexpr int MAXSEG = 128;

struct DevPlan {
    int d, P, Rmax, nmax, piv, kind, ising_id, nlotmax;
    int auxsm;             // doubles of dynamic shared memory reserved for the MVN matrix (0: read it from global)
    int stage;             // 1: evaluating kernels stage node/weight values of all pivots in shared memory
    int stage_max;         // doubles reserved for that staging area
    unsigned int* tickets; // [P] arrival counters of the last-CTA reductions
    const int* n;          // n[1..d]; n[0] = n[d+1] = 1
    const int* own;        // own[0..P]
    const double* par;     // nodes | weights | ...
    const double* aux;     // MVN: mu | inv_cov | denom
    int* Lidx; int* Ridx; const i64* offL; const i64* offR;
    int* vip;              // [(d+1)][Rmax][4]
    int* rk; int* rks;     // ranks now / at sweep start, index 0..d
    int* qsnap;            // [2][d+1] rank snapshots for the overlapped quadrature: the close of sweep s writes buffer s & 1, the
                           // quadrature of sweep s (second stream, beside the bond visits of sweep s+1) reads it; everything
                           // the quadrature reads below those ranks never changes once written
    unsigned int* btick;   // arrival counter of k_exchange_fused (its last CTA closes the sweep)
    double* arg; double* col; double* rowT; const i64* coreOff;   // coreOff[p], p = 1..d
    double* inv;           // [(d+1)][Rmax*Rmax]
    double* acol1; double* bcol1; double* arow1; double* brow1;   // [P][Rmax*nmax]
    int* lot;              // [P][4][nlotmax]
    double* lraw; double* lres;  // [P][nlotmax]
    Partial* part;         // [P][2][GMAX]
    VState* st;            // [P]
    VisitOut* out;         // [P]
    const double* quadw; const i64* quadOff;   // weights of core p at quadw + quadOff[p]
    double* ttqq;          // [(d+1)][Rmax*Rmax]
    double* chain;         // [P+1][Rmax*Rmax] partial products
    double* chain2;        // scratch, same size
    SweepOut* sweep_out;
    // asynchronous mode: nothing below needs the host during the sweeps
    int dev_lottery;       // 1: lottery on the device (built-in uniform stream), 0: host fills `lot`
    int maxnb, maxsweeps;
    Ctrl* ctrl;
    VisitOut* vlog;        // [maxsweeps][maxnb][P]
    SweepOut* slog;        // [maxsweeps + 1]
    int* rklog;            // [maxsweeps + 1][d + 1]
    // core blocks partitioned over processes (one process per GPU): this process runs virtual ranks v0 .. v0+nv-1 and
    // contracts / finalises cores c_lo .. c_hi (the dtt_lua / dtt_quad ownership of dmrgg.f90:1209-1257).  Metadata
    // (vip, Lidx, Ridx, rk, st) is replicated: foreign pivots are replayed from the all-gathered visit records.
    int v0, nv, nproc, prank, vper, c_lo, c_hi;
    unsigned long long* mb1_send; unsigned long long* mb1_recv;   // phase 1: per virtual rank [VisitOut x maxnb | VState]
    double* mb2_send; double* mb2_recv;                           // phase 2: per virtual rank [chain Rmax^2 | amax | neval | error | pad]
    double* nb_send_l; double* nb_recv_l;   // to/from the left neighbour process : send column slab [Rmax*nmax]; recv row [nmax*Rmax] | inv [Rmax^2]
    double* nb_send_r; double* nb_recv_r;   // to/from the right neighbour process: send row | inv; recv column slab
    // peer-memory exchange (one process per GPU, CUDA IPC): the four receive areas above live in ONE window per process;
    // peer_win[g] is rank g's window mapped into this process, the win_* are byte offsets inside a window.
    // flags[(phase - 1) * nproc + src] (phase 1, 2, 3 = final quadrature) carry the sequence number of the last push of `src`.
    char** peer_win; unsigned long long* win_flags;
    long long win_mb1, win_mb2, win_nbl, win_nbr, win_flg;
    double* ttqy;          // [(d+1)][Rmax*Rmax] contracted cores after d2_luar (incremental per-sweep quadrature, k_quad_inc)
    int* qext;             // [(d+1)][2] extents of ttqy/ttqq already computed
    // diagnostic timeline (ttc_set_timeline): every kernel stamps %globaltimer when its first CTA starts
    unsigned long long* tlog; int* tlog_n; int tlog_cap;
    // persistent sweep kernel (ttc_sweep.cuh): one mailbox per partition (inside the peer window when there are several processes)
    struct SweepMail* mail; long long win_mail;
    // several processes: what the neighbour processes push after their bond visits (double-buffered by sweep parity), and the
    // log / chain-product areas every process pushes into every window.  All are byte offsets inside a window.
    //   slab_l[par]: from the LEFT  process: new row of the shared core [nmax * Rmax] | L-table column [d ints] | packed-LU row [2 Rmax + 1]
    //   slab_r[par]: from the RIGHT process: new slice of the shared core [Rmax * nmax] | R-table column [d ints]
    long long win_sl, win_sr, slab_l_bytes, slab_r_bytes, win_vlog, win_rklog, win_chs;
    double* chainS;        // [maxsweeps][P + 1][Rmax^2] chain products of the per-sweep quadrature after the loop
    // VALUE tables of the persistent kernel: node value (and Ising weight) of every entry of Lidx / Ridx, same row structure with
    // an even leading dimension RT (16-byte rows), so that the evaluation inputs of a bond visit are plain contiguous blocks
    // that the TMA copies into shared memory (cp.async.bulk) instead of a gather through the index tables; parT = nodes | weights,
    // each padded to NT (even) entries.  Null when the persistent kernel is not in use.
    double* XLg; double* WLg; double* XRg; double* WRg; const double* parT; int RT, NT;
    int auxsm_p;           // doubles of shared memory the PERSISTENT kernel reserves for the MVN matrix (>= auxsm)
    int exp_mode;          // 0: platform exp; 1: the deterministic exp of include/ttc_detexp.h (parity mode, ttc_set_exp_mode)
};
__device__ __forceinline__ double plan_exp(const DevPlan& P, double x) { return P.exp_mode ? ttc_det_exp(x) : exp(x); }
__device__ __forceinline__ void tl_stamp(const DevPlan& P, int id) {
    if (P.tlog && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        int k = atomicAdd(P.tlog_n, 1);
        if (k < P.tlog_cap) { P.tlog[3 * k] = (unsigned long long)id; P.tlog[3 * k + 1] = t; P.tlog[3 * k + 2] = (unsigned long long)clock64(); }
    }
}
__host__ __device__ __forceinline__ int proc_v0(int P, int nproc, int g) { return (int)((long long)P * g / nproc); }
__device__ __forceinline__ bool own_vrank(const DevPlan& P, int v) { return v >= P.v0 && v < P.v0 + P.nv; }
// boundaries b (between virtual ranks b and b+1) that touch this process's virtual ranks: first one, and how many
__host__ __device__ __forceinline__ int first_boundary(const DevPlan& P) { return P.v0 > 0 ? P.v0 - 1 : 0; }
__host__ __device__ __forceinline__ int boundary_count(const DevPlan& P) {
    int last = P.v0 + P.nv - 1; if (last > P.P - 2) last = P.P - 2;
    return last - first_boundary(P) + 1;
}

__device__ __forceinline__ void tl_mark0(const DevPlan& P, int id) {     // diagnostic: phase stamps of the middle CTA
    if (P.tlog && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        int k = atomicAdd(P.tlog_n, 1);
        if (k < P.tlog_cap) { P.tlog[3 * k] = (unsigned long long)id; P.tlog[3 * k + 1] = t; P.tlog[3 * k + 2] = (unsigned long long)clock64(); }
    }
}
// ----------------------------------------------------------------------------
// bond-visit geometry (dmrgg.f90:329-331 and the rr/r snapshot of :325)
// ----------------------------------------------------------------------------
struct Dims { int active, p, r0, r1, r2, n1, n2; };

__device__ __forceinline__ Dims load_dims(const DevPlan& P, int v, int dir, int pp) {
    Dims D;
    int lo = P.own[v], hi = P.own[v + 1];
    D.active = (pp <= hi - lo);
    D.p = (dir == 1) ? lo + pp - 1 : hi - pp;
    if (!D.active) { D.p = lo; }
    int p = D.p;
    D.r0 = (p - 1 >= lo) ? P.rk[p - 1] : P.rks[p - 1];
    D.r1 = P.rk[p];
    D.r2 = (p + 1 <= hi - 1) ? P.rk[p + 1] : P.rks[p + 1];
    D.n1 = P.n[p];
    D.n2 = P.n[p + 1];
    return D;
}

// ----------------------------------------------------------------------------
// multi-index of one evaluation point: left table row | j | [k] | right table row
// (flat restatement of dmrgg_fun, dmrgg.f90:1053-1078)
// ----------------------------------------------------------------------------
struct PointSrc {
    const int* L; int nl; int i;       // positions 1..nl           : L[(pos-1)*Rmax + i-1]
    int j; int k; int hask;            // position nl+1 (and nl+2 when hask)
    const int* R; int q;               // positions nl+2+hask-1+.. : R[(pos')*Rmax + q-1]
    int Rmax;
    __device__ __forceinline__ int operator()(int pos) const {   // 1-based position -> 1-based mode index
        if (pos <= nl) return L[(i64)(pos - 1) * Rmax + (i - 1)];
        if (pos == nl + 1) return j;
        if (hask && pos == nl + 2) return k;
        return R[(i64)(pos - nl - 2 - hask) * Rmax + (q - 1)];
    }
};
__device__ __forceinline__ PointSrc bond_point(const DevPlan& P, int p, int i, int j, int k, int q) {
    PointSrc s;
    s.L = P.Lidx + P.offL[p - 1]; s.nl = p - 1; s.i = i; s.j = j; s.k = k; s.hask = 1;
    s.R = P.Ridx + P.offR[p + 1]; s.q = q; s.Rmax = P.Rmax;
    return s;
}
struct DiagSrc {   // wrapped diagonals of the initial search (dmrgg.f90:171-173)
    const int* n; int k; int s;
    __device__ __forceinline__ int operator()(int pos) const { return (k - 1 + s * (pos - 1)) % n[pos] + 1; }
};

// ----------------------------------------------------------------------------
// integrands.  A "value source" V supplies the node value x(pos) and the quadrature weight w(pos) of the mode index
// at 1-based position pos; the arithmetic below is the reference's, operation for operation.
// ----------------------------------------------------------------------------
// values straight from global memory through an index source (initial cross, fallback when staging does not fit)
template <class Src>
struct GlobalVals {
    Src s; const double* par; int nw;    // nw: offset of the weights inside par (Ising: n)
    __device__ __forceinline__ double x(int pos) const { return par[s(pos) - 1]; }
    __device__ __forceinline__ double w(int pos) const { return par[nw + s(pos) - 1]; }
    // all m positions at once: the index loads, then the value loads, are independent of each other (memory-level
    // parallelism instead of one dependent chain per accessor call)
    __device__ __forceinline__ void gather(int m, double* xv, double* wv, bool needw) const {
        for (int pos = 1; pos <= m; ++pos) {
            const int idx = s(pos) - 1;
            xv[pos - 1] = par[idx];
            if (needw) wv[pos - 1] = par[nw + idx];
        }
    }
};
// values staged in shared memory by stage_bond(): left table XL[pos][i], right table XR[pos][q], nodes/weights of the
// two free modes
struct StagedVals {
    const double* XL; const double* WL; int nl, rl, i;      // positions 1..nl, pivot i of rl
    double xj, wj, xk, wk; int hask;                        // position nl+1 (and nl+2)
    const double* XR; const double* WR; int rr, q;          // remaining positions, pivot q of rr
    __device__ __forceinline__ double x(int pos) const {
        if (pos <= nl) return XL[(pos - 1) * rl + (i - 1)];
        if (pos == nl + 1) return xj;
        if (hask && pos == nl + 2) return xk;
        return XR[(pos - nl - 2 - hask) * rr + (q - 1)];
    }
    __device__ __forceinline__ double w(int pos) const {
        if (pos <= nl) return WL[(pos - 1) * rl + (i - 1)];
        if (pos == nl + 1) return wj;
        if (hask && pos == nl + 2) return wk;
        return WR[(pos - nl - 2 - hask) * rr + (q - 1)];
    }
    // all m positions at once: three branch-free runs of independent shared-memory loads
    __device__ __forceinline__ void gather(int m, double* xv, double* wv, bool needw) const {
        const double* xl = XL + (i - 1); const double* wl = WL + (i - 1);
        if (needw) {
            for (int pos = 0; pos < nl; ++pos) { xv[pos] = xl[pos * rl]; wv[pos] = wl[pos * rl]; }
            xv[nl] = xj; wv[nl] = wj;
            int o = nl + 1;
            if (hask) { xv[o] = xk; wv[o] = wk; ++o; }
            const double* xr = XR + (q - 1); const double* wr = WR + (q - 1);
            for (int t = 0; o + t < m; ++t) { xv[o + t] = xr[t * rr]; wv[o + t] = wr[t * rr]; }
        } else {        // node values only (stdnorm, MVN): wv is not touched and may be null
            for (int pos = 0; pos < nl; ++pos) xv[pos] = xl[pos * rl];
            xv[nl] = xj;
            int o = nl + 1;
            if (hask) { xv[o] = xk; ++o; }
            const double* xr = XR + (q - 1);
            for (int t = 0; o + t < m; ++t) xv[o + t] = xr[t * rr];
        }
    }
};

// test_crs_ising.f90:176-218.  Pure + - * / : bit-reproducible.
// The node values / weights of the m positions are gathered first (independent loads), then the reference's
// recurrences run on the arrays: same operations in the same order, without a branchy accessor inside every step.
template <class V>
__device__ __noinline__ double eval_ising(const DevPlan& P, const V& v) {
    const int m = P.d;
    const int id = P.ising_id;
    double a = 0.0, b = 0.0, f;
    if (m <= MAXD_LOCAL) {
        double x[MAXD_LOCAL], wq[MAXD_LOCAL];
        v.gather(m, x, wq, true);
        if (id == 2 || id == 3) {
            a = 1.0;
#pragma unroll 1
            for (int i = 0; i <= m; ++i) {
                double uij = 1.0;
#pragma unroll 2
                for (int j = i + 1; j <= m; ++j) {
                    uij = uij * x[j - 1];
                    double t = (uij - 1.0) / (uij + 1.0);
                    a = a * (t * t);
                }
            }
        }
        if (id == 1 || id == 2) {
            double vv = 1.0, w = 1.0, vk = 1.0, wk = 1.0;
#pragma unroll 4
            for (int i = 1; i <= m; ++i) {
                vk = vk * x[m - i];
                wk = wk * x[i - 1];
                vv = vv + vk;
                w = w + wk;
            }
            b = 1.0 / (vv * w);
        }
        if (id == 1) f = 2 * b;
        else if (id == 2) f = 2 * a * b;
        else f = 2 * a;
#pragma unroll 4
        for (int i = 0; i < m; ++i) f = f * wq[i];
        return f;
    }
    if (id == 2 || id == 3) {
        a = 1.0;
        for (int i = 0; i <= m; ++i) {
            double uij = 1.0;
            for (int j = i + 1; j <= m; ++j) {
                uij = uij * v.x(j);
                double t = (uij - 1.0) / (uij + 1.0);
                a = a * (t * t);
            }
        }
    }
    if (id == 1 || id == 2) {
        double vv = 1.0, w = 1.0, vk = 1.0, wk = 1.0;
        for (int i = 1; i <= m; ++i) {
            vk = vk * v.x(m - i + 1);
            wk = wk * v.x(i);
            vv = vv + vk;
            w = w + wk;
        }
        b = 1.0 / (vv * w);
    }
    if (id == 1) f = 2 * b;
    else if (id == 2) f = 2 * a * b;
    else f = 2 * a;
    for (int i = 1; i <= m; ++i) f = f * v.w(i);
    return f;
}
// test_crs_stdnorm.f90:154-170
template <class V>
__device__ __noinline__ double eval_stdnorm(const DevPlan& P, const V& v) {
    double sum = 0.0;
    if (P.d <= MAXD_LOCAL) {
        double x[MAXD_LOCAL];
        v.gather(P.d, x, nullptr, false);
        for (int i = 0; i < P.d; ++i) sum = sum + x[i] * x[i];
        return plan_exp(P, -sum);
    }
    for (int i = 1; i <= P.d; ++i) { double x = v.x(i); sum = sum + x * x; }
    return plan_exp(P, -sum);
}
// lib/mvn_pdf.f90:63-83 (through test_crs_mvn.f90:156-172); A = inv_cov column-major, staged by the caller
template <class V>
__device__ __noinline__ double eval_mvn(const DevPlan& P, const V& v, const double* __restrict__ A /*d*d*/) {
    const int m = P.d;
    const double* mu = P.aux;
    const double denom = P.aux[m + (i64)m * m];
    double e = 0.0;
    if (m <= MAXD_LOCAL) {
        double diff[MAXD_LOCAL];
        v.gather(m, diff, nullptr, false);
        for (int i = 0; i < m; ++i) diff[i] = diff[i] - mu[i];
        for (int i = 0; i < m; ++i) {
            const double di = diff[i];
            for (int j = 0; j < m; ++j) e = e + di * A[i + (i64)j * m] * diff[j];
        }
    } else {
        for (int i = 0; i < m; ++i) {
            const double di = v.x(i + 1) - mu[i];
            for (int j = 0; j < m; ++j) e = e + di * A[i + (i64)j * m] * (v.x(j + 1) - mu[j]);
        }
    }
    return plan_exp(P, -0.5 * e) / denom;
}

// The same arithmetic with diff[] held in REGISTERS for the inner loop.  eval_mvn above keeps diff[] in local memory: at
// d = 64 that is 512 B per thread, the resident threads' copies (256 KB per SM) do not fit the L1 next to the staged
// matrix, and every one of the d*d terms re-reads diff[j] from L2 (timeline of config E: 300 us per evaluation).  Here the
// inner loop over j is fully unrolled over M >= d slots in blocks of 8 (d a multiple of 8), so the only memory operand of
// a term is the matrix entry, which all lanes share.  Needs ~2 M + 40 registers: kernels that call it must not cap the
// register count below that (k_visits<MVN> is compiled for one CTA per SM).  Same operations in the same order.
template <int M, class V>
__device__ __noinline__ double eval_mvn_reg(const DevPlan& P, const V& v, const double* __restrict__ A /*d*d*/) {
    const int m = P.d;                                   // a multiple of 8, M - 16 < m <= M
    const double* mu = P.aux;
    const double denom = P.aux[m + (i64)m * m];
    double x[MAXD_LOCAL];
    v.gather(m, x, nullptr, false);
    double dr[M];
#pragma unroll
    for (int j = 0; j < M; ++j) dr[j] = (j < m) ? x[j] - mu[j] : 0.0;
    double e = 0.0;
    for (int i = 0; i < m; ++i) {
        const double di = x[i] - mu[i];
        const double* Ai = A + i;
#pragma unroll
        for (int jb = 0; jb < M; jb += 8) {
            if (jb < m) {                                // whole blocks of 8: the matrix entries load ahead of the chain
                double a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = Ai[(i64)(jb + u) * m];
#pragma unroll
                for (int u = 0; u < 8; ++u) e = e + di * a[u] * dr[jb + u];
            }
        }
    }
    return plan_exp(P, -0.5 * e) / denom;
}
template <class V>
__device__ __forceinline__ double eval_mvn_fast(const DevPlan& P, const V& v, const double* A) {
    const int m = P.d;
    if (m <= 16 || m > 64 || (m & 7)) return eval_mvn(P, v, A);     // small d: the local copy stays in L1
    if (m <= 32) return eval_mvn_reg<32>(P, v, A);
    if (m <= 48) return eval_mvn_reg<48>(P, v, A);
    return eval_mvn_reg<64>(P, v, A);
}

// COS-method coefficient of a Gaussian density (lib/coefficients.f90:33-65 with lib/funcs.f90:8-26, lib/s_vectors.f90:7-29):
//   f = 2/(b-a)^d * sum over the 2^(d-1) sign vectors s (s_1 = 1) of Re[ exp(-i a sum(t)) * exp(i t.mu - t.Sigma.t / 2) ],
//   t_j = pi s_j (ind_j - 1) / (b - a).   aux = mu(d) | Sigma(d,d) column-major | a | b;  the "node" of mode index k is k - 1.
// x ** n with an integer n is libgcc's __powidf2 (square and multiply), matmul(Sigma, t) accumulates column by column.
__host__ __device__ __forceinline__ double powi_gcc(double x, int m) {
    unsigned int n = m < 0 ? (unsigned)(-m) : (unsigned)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n % 2) y = y * x; }
    return m < 0 ? 1.0 / y : y;
}
constexpr int COS_MAXD = 24;
template <class V>
__device__ __noinline__ double eval_coscoef(const DevPlan& P, const V& v) {
    const int m = P.d;
    if (m > COS_MAXD) return nan("");
    const double* mu = P.aux; const double* sg = P.aux + m;
    const double lower = P.aux[m + (i64)m * m], upper = P.aux[m + (i64)m * m + 1];
    const double pi = 3.14159265358979323846;
    double k[COS_MAXD], t[COS_MAXD], y[COS_MAXD];
    v.gather(m, k, y, false);
    const double oob = 1 / (upper - lower);
    const double factor = 2.0 * powi_gcc(oob, m);
    double real_sum = 0.0;
    const unsigned ns = 1u << (m - 1);
    for (unsigned i = 0; i < ns; ++i) {
        for (int j = 0; j < m; ++j) {
            const int sj = (j == 0) ? 1 : (((i >> (j - 1)) & 1u) ? -1 : 1);
            t[j] = (((pi * (double)sj) * k[j]) * oob);
        }
        double dot_mu = 0.0, st = 0.0;
        for (int j = 0; j < m; ++j) { dot_mu = dot_mu + t[j] * mu[j]; st = st + t[j]; }
        for (int a = 0; a < m; ++a) y[a] = 0.0;
        for (int b = 0; b < m; ++b) for (int a = 0; a < m; ++a) y[a] = y[a] + sg[a + (i64)b * m] * t[b];
        double quad = 0.0;
        for (int a = 0; a < m; ++a) quad = quad + y[a] * t[a];
        const double E = exp(-0.5 * quad);
        double s2, c2, s1, c1;
        sincos(dot_mu, &s2, &c2);
        sincos(-lower * st, &s1, &c1);
        real_sum = real_sum + (c1 * (E * c2) - s1 * (E * s2));
    }
    return factor * real_sum;
}

// Stage the MVN matrix into shared memory when it fits; returns the pointer the integrand should read.
template <int KIND>
__device__ __forceinline__ const double* stage_aux(const DevPlan& P, double* smem) {
    if (KIND != KIND_MVN) return nullptr;
    const int m = P.d;
    const double* A = P.aux + m;
    if ((i64)m * m <= P.auxsm) {
        for (int x = threadIdx.x; x < m * m; x += blockDim.x) smem[x] = A[x];
        __syncthreads();
        return smem;
    }
    return A;
}
template <int KIND, class V>
__device__ __forceinline__ double eval_point(const DevPlan& P, const V& v, const double* A) {
    if (KIND == KIND_ISING || KIND == KIND_ISINGC) return eval_ising(P, v);
    if (KIND == KIND_STDNORM) return eval_stdnorm(P, v);
    if (KIND == KIND_COSCOEF) return eval_coscoef(P, v);
    return eval_mvn(P, v, A);
}
// for kernels without a tight register cap (see eval_mvn_reg)
template <int KIND, class V>
__device__ __forceinline__ double eval_point_wide(const DevPlan& P, const V& v, const double* A) {
    if (KIND == KIND_MVN) return eval_mvn_fast(P, v, A);
    return eval_point<KIND>(P, v, A);
}
template <int KIND, class Src>
__device__ __forceinline__ double eval_src(const DevPlan& P, const Src& s, const double* A) {
    GlobalVals<Src> v{s, P.par, P.n[1]};
    return eval_point<KIND>(P, v, A);
}

// ----------------------------------------------------------------------------
// per-CTA staging of everything an evaluation reads: node/weight values of the two free modes and of every pivot of
// the left table (bond pl, nl positions, rl pivots) and of the right table (bond pr, nr positions, rr pivots).
// Shared-memory layout (doubles): NX[n1] NW[n1] NX2[n2] NW2[n2] XL[nl*rl] WL[nl*rl] XR[nr*rr] WR[nr*rr]
// ----------------------------------------------------------------------------
struct Stage {
    const double *NX, *NW, *NX2, *NW2, *XL, *WL, *XR, *WR;
    int nl, rl, nr, rr, hask;
    __device__ __forceinline__ StagedVals point(int i, int j, int k, int q) const {
        StagedVals v;
        v.XL = XL; v.WL = WL; v.nl = nl; v.rl = rl; v.i = i;
        v.xj = NX[j - 1]; v.wj = NW[j - 1];
        v.hask = hask;
        v.xk = hask ? NX2[k - 1] : 0.0; v.wk = hask ? NW2[k - 1] : 0.0;
        v.XR = XR; v.WR = WR; v.rr = rr; v.q = q;
        return v;
    }
};
__host__ __device__ __forceinline__ i64 stage_doubles(int n1, int n2, int nl, int rl, int nr, int rr) {
    return 2LL * n1 + 2LL * n2 + 2LL * nl * rl + 2LL * nr * rr;
}
// pl: bond whose left multi-indices feed positions 1..nl (nl = pl); pr: bond whose right multi-indices feed the tail.
// c1/c2: the cores of the free modes (c2 = 0 when there is a single free mode).
__device__ __forceinline__ Stage stage_bond(const DevPlan& P, double* sm, int pl, int rl, int c1, int c2, int pr, int rr) {
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
        int idx = L[(i64)pos * P.Rmax + t];
        XL[x] = P.par[idx - 1]; WL[x] = hasw ? P.par[nwoff + idx - 1] : 0.0;
    }
    const int* R = P.Ridx + P.offR[pr];
    for (int x = threadIdx.x; x < nr * rr; x += blockDim.x) {
        int pos = x / rr, t = x - pos * rr;
        int idx = R[(i64)pos * P.Rmax + t];
        XR[x] = P.par[idx - 1]; WR[x] = hasw ? P.par[nwoff + idx - 1] : 0.0;
    }
    __syncthreads();
    S.NX = NX; S.NW = NW; S.NX2 = NX2; S.NW2 = NW2; S.XL = XL; S.WL = WL; S.XR = XR; S.WR = WR;
    S.nl = nl; S.rl = rl; S.nr = nr; S.rr = rr; S.hask = c2 ? 1 : 0;
    return S;
}

// ----------------------------------------------------------------------------
// first-index argmax of |x| (netlib idamax: strict '>' while scanning upwards, NaN never wins)
// ----------------------------------------------------------------------------
__device__ __forceinline__ void amax_take(Partial& a, double val, i64 idx) {
    double av = fabs(val);
    if (av > a.absv) { a.absv = av; a.val = val; a.idx = idx; }
}
__device__ __forceinline__ void amax_merge(Partial& a, const Partial& b) {
    if (b.absv > a.absv || (b.absv == a.absv && b.idx < a.idx)) a = b;
}
__device__ __forceinline__ Partial amax_init() { Partial p; p.absv = -1.0; p.val = 0.0; p.idx = 0x7fffffffffffffffLL; return p; }
__device__ __forceinline__ Partial amax_warp(Partial a) {
    for (int o = 16; o > 0; o >>= 1) {
        Partial b;
        b.absv = __shfl_down_sync(0xffffffffu, a.absv, o);
        b.val = __shfl_down_sync(0xffffffffu, a.val, o);
        b.idx = __shfl_down_sync(0xffffffffu, a.idx, o);
        amax_merge(a, b);
    }
    return a;
}
// block reduce; result valid in thread 0.  sh must hold 32 Partials.
__device__ __forceinline__ Partial amax_block(Partial a, Partial* sh) {
    a = amax_warp(a);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = a;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        a = (lane < nw) ? sh[lane] : amax_init();
        a = amax_warp(a);
    }
    return a;
}

// ----------------------------------------------------------------------------
// device-side lottery (rnd.f90:105-144 + dmrgg.f90:425-452), bit-exact.
//
// The reference draws a cell from cumulative weights pcol(i) = pcol(i-1) + |w_i|/scol accumulated SEQUENTIALLY in
// double precision.  Weights are 1 except 0 at the existing pivots, so pcol(i) = T[c(i)] with c(i) = number of non-zero
// weights among the first i cells and T[c] = fl(T[c-1] + delta), delta = fl(1/scol).  T is strictly increasing, hence
// find_d's bisection returns the unique cell whose count c* satisfies T[c*-1] <= y < T[c*] whatever path it takes.
// T is evaluated in closed form: inside one binade the rounded increment is a constant integer number of ulps
// (round-to-nearest; a tie settles into a constant after one step), so T is piecewise linear in exact integer
// arithmetic; the (few) steps that cross a binade are done with a real floating-point addition.
// ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double stream_uniform(unsigned long long seed, int vrank, unsigned long long k) {
    const unsigned long long G = 0x9E3779B97F4A7C15ULL;
    unsigned long long base = mix64(seed + G * (unsigned long long)(vrank + 1));
    unsigned long long z = mix64(base + G * (k + 1));
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
__host__ __device__ __forceinline__ long long dbl_bits(double x) {
#ifdef __CUDA_ARCH__
    return __double_as_longlong(x);
#else
    long long b; memcpy(&b, &x, sizeof b); return b;
#endif
}
// segments cover counts 1..scol; returns the number of segments
__host__ __device__ __noinline__ int build_segments(int scol, LotSeg* seg) {
    const double delta = 1.0 / (double)scol;
    const long long db = dbl_bits(delta);
    const int Ed = (int)((db >> 52) & 0x7ff) - 1023;
    const long long Md = (db & ((1LL << 52) - 1)) | (1LL << 52);
    int c = 1, ns = 0;
    double T = delta;
    while (true) {
        const long long tb = dbl_bits(T);
        const int E = (int)((tb >> 52) & 0x7ff) - 1023;
        const long long M = (tb & ((1LL << 52) - 1)) | (1LL << 52);
        const int sh = E - Ed;
        const long long q = Md >> sh;
        const long long rem = Md & ((1LL << sh) - 1);
        const long long half = sh ? (1LL << (sh - 1)) : 0;
        long long k = q;
        bool single = false;
        if (sh > 0) {
            if (rem > half) k = q + 1;
            else if (rem == half) { if (M & 1) single = true; else k = q + (q & 1); }   // ties to even
        }
        // a step from mantissa m is an in-binade step iff m + q <= 2^53 - 1 (exact sum stays below the next power of two)
        const long long lim = (1LL << 53) - 1 - q - M;
        long long J = (single || lim < 0) ? 0 : lim / k + 1;
        if (J > scol - c) J = scol - c;
        if (ns < MAXSEG) { seg[ns].M = M; seg[ns].k = k; seg[ns].c0 = c; seg[ns].J = (int)J; seg[ns].E = E; ++ns; }
        c += (int)J;
        if (c >= scol) break;
        T = scalbn((double)(M + J * k), E - 52);
        T = T + delta;              // binade-crossing (or odd-tie) step: hardware rounding
        c += 1;
    }
    return ns;
}
__host__ __device__ __noinline__ double lot_T(const LotSeg* seg, int ns, int c) {   // T[c], 1 <= c <= scol
    int lo = 0, hi = ns - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (seg[mid].c0 <= c) lo = mid; else hi = mid - 1; }
    const LotSeg& g = seg[lo];
    return scalbn((double)(g.M + (long long)(c - g.c0) * g.k), g.E - 52);
}
// one draw: 1-based cell index in 1..m (zeros = sorted distinct 1-based zero-weight cells)
__host__ __device__ __noinline__ int lot_draw(const LotSeg* seg, int ns, int scol, int m, const int* zeros, int nz, double y) {
    if (!(y < lot_T(seg, ns, scol))) return m;          // x(n) <= y  ->  n = m+1, clamped to m (rnd.f90:122)
    // smallest count c with y < T[c].  T[c] is c/scol up to a few ulps, so start next to y*scol and walk (1-2 steps)
    // instead of bisecting: every lot_T is a dependent chain of shared-memory loads.
    int lo = (int)(y * (double)scol) + 1;
    if (lo < 1) lo = 1;
    if (lo > scol) lo = scol;
    while (lo < scol && !(y < lot_T(seg, ns, lo))) ++lo;
    while (lo > 1 && y < lot_T(seg, ns, lo - 1)) --lo;
    int sidx = lo;                                       // the lo-th cell of non-zero weight
    for (int z = 0; z < nz; ++z) { if (zeros[z] <= sidx) ++sidx; else break; }
    return sidx;
}
// The same draw WITHOUT the segment table (used by the cluster kernel): T[c] = c*delta up to the accumulated rounding of c
// additions, |T[c] - c*delta| <= c * 2^-53 * T[c] < (c + 4) * 2^-52, so the comparison y < T[c] is decided by one
// multiplication unless y lies inside that window (probability ~1e-9 per draw); only then T[c] is formed exactly as the
// reference forms it, by c sequential additions (rnd.f90:118-119).
__host__ __device__ __forceinline__ bool lot_lt(double y, int c, double delta) {      // y < T[c] ?
    const double a = (double)c * delta;
    const double tol = (double)(c + 4) * 2.220446049250313e-16;
    if (y < a - tol) return true;
    if (y >= a + tol) return false;
    double T = 0.0;
    for (int i = 0; i < c; ++i) T = T + delta;
    return y < T;
}
__host__ __device__ __noinline__ int lot_draw_fast(int scol, int m, const int* zeros, int nz, double y) {
    const double delta = 1.0 / (double)scol;
    if (!lot_lt(y, scol, delta)) return m;              // x(n) <= y  ->  n = m+1, clamped to m (rnd.f90:122)
    int lo = (int)(y * (double)scol) + 1;
    if (lo < 1) lo = 1;
    if (lo > scol) lo = scol;
    while (lo < scol && !lot_lt(y, lo, delta)) ++lo;
    while (lo > 1 && lot_lt(y, lo - 1, delta)) --lo;
    // the lo-th cell of non-zero weight = lo + (number of zero cells before it).  Walking the sorted distinct zeros
    // (`if (zeros[z] <= sidx) ++sidx; else break;`) takes zero z exactly when zeros[z] - z <= lo, and zeros[z] - z never
    // decreases, so the walk equals this count -- with independent loads instead of a dependent chain.
    int k = 0;
    for (int z = 0; z < nz; ++z) k += (zeros[z] - z <= lo) ? 1 : 0;
    return lo + k;
}
// sorted distinct zero-weight cells (1-based) of one side of bond p; thread-parallel rank sort over the r1 pivots.
// tmp, zeros: shared int[>= r1]; *nz: shared int.  side 0: (i,j) with stride r0; side 1: (k,q) with stride n2.
__device__ __forceinline__ void lot_zeros(const int* vip_p, int r1, int side, int stride, int* tmp, int* zeros, int* nz) {
    for (int t = threadIdx.x; t < r1; t += blockDim.x)
        zeros[t] = (vip_p[4 * t + 2 * side] - 1) + stride * (vip_p[4 * t + 2 * side + 1] - 1) + 1;
    __syncthreads();
    for (int t = threadIdx.x; t < r1; t += blockDim.x) {          // stable rank sort (duplicates kept)
        int me = zeros[t], rank = 0;
        for (int u = 0; u < r1; ++u) { int o = zeros[u]; rank += (o < me) || (o == me && u < t); }
        tmp[rank] = me;
    }
    __syncthreads();
    int keep = 0, pos = 0;
    for (int t = threadIdx.x; t < r1; t += blockDim.x) {          // compact the distinct values (r1 <= blockDim.x expected, loop anyway)
        keep = (t == 0) || (tmp[t] != tmp[t - 1]);
        pos = 0;
        for (int u = 1; u <= t; ++u) pos += (tmp[u] != tmp[u - 1]);
        if (keep) zeros[pos] = tmp[t];
        if (t == r1 - 1) *nz = pos + 1;
    }
    __syncthreads();
}

// both sides in one pass (three block barriers instead of six).  tmp: shared int[2*r1]
__device__ __forceinline__ void lot_zeros2(const int* vip_p, int r1, int r0, int n2, int* tmp, int* zc, int* zr, int* nz /*[2]*/) {
    int* tc = tmp; int* tr = tmp + r1;
    for (int t = threadIdx.x; t < r1; t += blockDim.x) {
        const int4 vp = *reinterpret_cast<const int4*>(vip_p + 4 * t);
        zc[t] = (vp.x - 1) + r0 * (vp.y - 1) + 1;
        zr[t] = (vp.z - 1) + n2 * (vp.w - 1) + 1;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < r1; t += blockDim.x) {
        const int mc = zc[t], mr = zr[t];
        int rc = 0, rr = 0;
        for (int u = 0; u < r1; ++u) {
            const int oc = zc[u], orr = zr[u];
            rc += (oc < mc) || (oc == mc && u < t);
            rr += (orr < mr) || (orr == mr && u < t);
        }
        tc[rc] = mc; tr[rr] = mr;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < r1; t += blockDim.x) {
        const int kc = (t == 0) || (tc[t] != tc[t - 1]), kr = (t == 0) || (tr[t] != tr[t - 1]);
        int pc = 0, pr = 0;
        for (int u = 1; u <= t; ++u) { pc += (tc[u] != tc[u - 1]); pr += (tr[u] != tr[u - 1]); }
        if (kc) zc[pc] = tc[t];
        if (kr) zr[pr] = tr[t];
        if (t == r1 - 1) { nz[0] = pc + 1; nz[1] = pr + 1; }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------
// shared helpers of the evaluating kernels
// ----------------------------------------------------------------------------
// grid-wide first-index argmax without a second launch: every CTA publishes its partials, the last CTA to arrive
// (atomic ticket) folds them and updates the visit state (the classic threadfence reduction).
// Returns true in the last CTA; `raw`/`res` then hold the folded results in thread 0.
__device__ __forceinline__ bool fold_partials(const DevPlan& P, int v, Partial& raw, Partial& res, Partial* shp) {
    __shared__ int s_last;
    __threadfence();                 // this thread's fiber / lottery stores, before the CTA announces itself
    raw = amax_block(raw, shp);
    res = amax_block(res, shp);
    Partial* part = P.part + (i64)v * 2 * GMAX;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = raw;
        part[GMAX + blockIdx.x] = res;
        __threadfence();
        unsigned t = atomicAdd(P.tickets + v, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    const volatile Partial* vp = part;
    Partial a = amax_init(), b2 = amax_init();
    for (int x = threadIdx.x; x < (int)gridDim.x; x += blockDim.x) {
        Partial t1, t2;
        t1.absv = vp[x].absv; t1.val = vp[x].val; t1.idx = vp[x].idx;
        t2.absv = vp[GMAX + x].absv; t2.val = vp[GMAX + x].val; t2.idx = vp[GMAX + x].idx;
        amax_merge(a, t1); amax_merge(b2, t2);
    }
    raw = amax_block(a, shp);
    res = amax_block(b2, shp);
    if (threadIdx.x == 0) {
        P.tickets[v] = 0;
        if (P.tlog && v == P.v0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            int k = atomicAdd(P.tlog_n, 1);
            if (k < P.tlog_cap) { P.tlog[3 * k] = 100ULL; P.tlog[3 * k + 1] = t; P.tlog[3 * k + 2] = (unsigned long long)clock64(); }
        }
    }
    return true;
}
// residuals in the reference's orders, with the factor loads issued in batches (they do not depend on the sum)
constexpr int RU = 8;
// dgemv 'n' / dgemm order: res = f; res += (-x_s) * a_s, s ascending   (a_s = base[s*stride])
__device__ __forceinline__ double resid_axpy(double f, const double* base, i64 stride, const double* xs, int r) {
    double res = f;
    int s0 = 0;
    for (; s0 + RU <= r; s0 += RU) {
        double a[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) a[u] = base[(s0 + u) * stride];
#pragma unroll
        for (int u = 0; u < RU; ++u) res = res + (-xs[s0 + u]) * a[u];
    }
    for (; s0 < r; ++s0) res = res + (-xs[s0]) * base[s0 * stride];
    return res;
}
// dgemv 't' order: t = sum_s a_s * x_s from 0, ascending; result f + (-t)
__device__ __forceinline__ double resid_dot(double f, const double* base, i64 stride, const double* xs, int r) {
    double t = 0.0;
    int s0 = 0;
    for (; s0 + RU <= r; s0 += RU) {
        double a[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) a[u] = base[(s0 + u) * stride];
#pragma unroll
        for (int u = 0; u < RU; ++u) t = t + a[u] * xs[s0 + u];
    }
    for (; s0 < r; ++s0) t = t + base[s0 * stride] * xs[s0];
    return f + (-t);
}
// ddot(r, col(i,j,1), ldc, row(1,k,q), 1) (dmrgg.f90:474): t = sum_s c_s * r_s from 0; result f - t
__device__ __forceinline__ double resid_ddot2(double f, const double* c, i64 cs, const double* r, i64 rs, int r1) {
    double t = 0.0;
    int s0 = 0;
    for (; s0 + RU <= r1; s0 += RU) {
        double a[RU], b[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) { a[u] = c[(s0 + u) * cs]; b[u] = r[(s0 + u) * rs]; }
#pragma unroll
        for (int u = 0; u < RU; ++u) t = t + a[u] * b[u];
    }
    for (; s0 < r1; ++s0) t = t + c[s0 * cs] * r[s0 * rs];
    return f - t;
}
template <int KIND>
__device__ __forceinline__ double eval_bond(const DevPlan& P, const Stage& S, int p, int i, int j, int k, int q, const double* A) {
    if (P.stage) { StagedVals v = S.point(i, j, k, q); return eval_point<KIND>(P, v, A); }
    return eval_src<KIND>(P, bond_point(P, p, i, j, k, q), A);
}

// ----------------------------------------------------------------------------
// K1: lottery candidates (dmrgg.f90:425-484): draw (device lottery), evaluate, residual by sequential ddot,
// two first-index argmaxes, and the scalar bookkeeping of dmrgg.f90:465-490 in the last CTA.
// dynamic smem: A[auxsm] | stage | ints[3*Rmax]
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_lot(DevPlan P, int dir, int pp) {
    tl_stamp(P, 0);
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const double* A = stage_aux<KIND>(P, smem);
    double* stg = smem + P.auxsm;
    Stage S;
    if (P.stage) S = stage_bond(P, stg, D.p - 1, D.r0, D.p, D.p + 1, D.p + 1, D.r2);
    const int nlot = D.r0 + D.n1 + D.n2 + D.r2;
    int* lot = P.lot + (i64)v * 4 * P.nlotmax;
    if (P.dev_lottery) {
        // every CTA rebuilds the (tiny) cumulative-weight description; each thread then draws its own candidates
        __shared__ LotSeg seg[2][MAXSEG];
        __shared__ int s_ns[2], s_nz[2];
        int* ibuf = (int*)(stg + P.stage_max);         // tmp[Rmax] | zeros_col[Rmax] | zeros_row[Rmax]
        int* tmp = ibuf; int* zc = ibuf + P.Rmax; int* zr = ibuf + 2 * P.Rmax;
        const int* vip_p = P.vip + (i64)D.p * P.Rmax * 4;
        const int m = D.r0 * D.n1, n = D.n2 * D.r2;
        lot_zeros(vip_p, D.r1, 0, D.r0, tmp, zc, &s_nz[0]);
        lot_zeros(vip_p, D.r1, 1, D.n2, tmp, zr, &s_nz[1]);
        if (threadIdx.x == 0) s_ns[0] = build_segments(m - s_nz[0], seg[0]);
        if (threadIdx.x == 32) s_ns[1] = build_segments(n - s_nz[1], seg[1]);
        __syncthreads();
        const unsigned long long k0 = P.st[v].rng_k;
        const unsigned long long seed = P.ctrl->seed;
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nlot; x += gridDim.x * blockDim.x) {
            double uc = stream_uniform(seed, v, k0 + (unsigned long long)x);
            double ur = stream_uniform(seed, v, k0 + (unsigned long long)(nlot + x));
            int c = lot_draw(seg[0], s_ns[0], m - s_nz[0], m, zc, s_nz[0], uc);
            int w = lot_draw(seg[1], s_ns[1], n - s_nz[1], n, zr, s_nz[1], ur);
            lot[x] = (c - 1) % D.r0 + 1;
            lot[P.nlotmax + x] = (c - 1) / D.r0 + 1;
            lot[2 * P.nlotmax + x] = (w - 1) % D.n2 + 1;
            lot[3 * P.nlotmax + x] = (w - 1) / D.n2 + 1;
        }
        // (each thread reads back only the entries it wrote itself)
    }
    const double* colp = P.col + P.coreOff[D.p];
    const double* rowp = P.rowT + P.coreOff[D.p + 1];
    const i64 cs = (i64)P.Rmax * D.n1;       // stride of s in col(i,j,s)
    const i64 rs = (i64)D.n2 * P.Rmax;       // stride of s in rowT(s,k,q)
    Partial braw = amax_init(), bres = amax_init();
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nlot; x += gridDim.x * blockDim.x) {
        int i = lot[x], j = lot[P.nlotmax + x], k = lot[2 * P.nlotmax + x], q = lot[3 * P.nlotmax + x];
        double f = eval_bond<KIND>(P, S, D.p, i, j, k, q, A);
        double res = resid_ddot2(f, colp + (i - 1) + (i64)P.Rmax * (j - 1), cs, rowp + (k - 1) + (i64)D.n2 * (q - 1), rs, D.r1);
        P.lres[(i64)v * P.nlotmax + x] = res;
        amax_take(braw, f, x);
        amax_take(bres, res, x);
    }
    if (!fold_partials(P, v, braw, bres, shp)) return;
    if (threadIdx.x == 0) {
        VState& St = P.st[v];
        St.amax = fmax(St.amax, braw.absv);
        int x = (int)bres.idx;
        const volatile int* vl = lot;
        St.ii = vl[x]; St.jj = vl[P.nlotmax + x]; St.kk = vl[2 * P.nlotmax + x]; St.qq = vl[3 * P.nlotmax + x];
        St.pivot = bres.val;
        St.done = 0; St.havecol = 0; St.haverow = 0; St.crs = 0; St.upd = 0;
        St.neval += nlot;
        St.rng_k += 2ULL * (unsigned long long)nlot;     // one random_number(d(npnt,2)) call (rnd.f90:120)
    }
}

// ----------------------------------------------------------------------------
// K2: cross fibers of the rook search (dmrgg.f90:519-581) fused with their residuals and the scalar bookkeeping
//   column fiber: acol1(i,j) = f(i,j,kk,qq); bcol1 = acol1 - col(p)(:,:,1:r) * row(p+1)(1:r,kk,qq)   [dgemv 'n' order]
//   row fiber   : arow1(k,q) = f(ii,jj,k,q); brow1 = arow1 - row(p+1)(1:r,:,:)^T col(p)(ii,jj,1:r)   [dgemv 't' order]
// mode 0: rook step (skipped when the visit is already `done`); mode 1: piv = 0 (no argmax, dmrgg.f90:492-513);
// mode 2: piv = -1 (the fibers are slices of the superblock: nothing to account)
// dynamic smem: A[auxsm] | xs[Rmax] | stage
// ----------------------------------------------------------------------------
template <int KIND, int ISROW>
__global__ void k_fiber(DevPlan P, int dir, int pp, int mode) {
    tl_stamp(P, 1);
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const VState& St0 = P.st[v];
    if (mode == 0 && St0.done) return;
    const int ii = St0.ii, jj = St0.jj, kk = St0.kk, qq = St0.qq;
    const double* A = stage_aux<KIND>(P, smem);
    double* xs = smem + P.auxsm;               // [Rmax] coefficient vector of the residual
    double* stg = xs + P.Rmax;
    const double* colp = P.col + P.coreOff[D.p];
    const double* rowp = P.rowT + P.coreOff[D.p + 1];
    const i64 cs = (i64)P.Rmax * D.n1;
    const i64 rs = (i64)D.n2 * P.Rmax;
    for (int s = threadIdx.x; s < D.r1; s += blockDim.x)
        xs[s] = ISROW ? colp[(ii - 1) + (i64)P.Rmax * (jj - 1) + s * cs] : rowp[(kk - 1) + (i64)D.n2 * (qq - 1) + s * rs];
    Stage S;
    if (P.stage) S = stage_bond(P, stg, D.p - 1, D.r0, D.p, D.p + 1, D.p + 1, D.r2);
    else __syncthreads();
    const int count = ISROW ? D.n2 * D.r2 : D.r0 * D.n1;
    double* fa = (ISROW ? P.arow1 : P.acol1) + (i64)v * P.Rmax * P.nmax;
    double* fb = (ISROW ? P.brow1 : P.bcol1) + (i64)v * P.Rmax * P.nmax;
    Partial braw = amax_init(), bres = amax_init();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        double f, res;
        if (!ISROW) {
            int j = e / D.r0 + 1, i = e % D.r0 + 1;
            f = eval_bond<KIND>(P, S, D.p, i, j, kk, qq, A);
            res = resid_axpy(f, colp + (i - 1) + (i64)P.Rmax * (j - 1), cs, xs, D.r1);
        } else {
            int q = e / D.n2 + 1, k = e % D.n2 + 1;
            f = eval_bond<KIND>(P, S, D.p, ii, jj, k, q, A);
            res = resid_dot(f, rowp + (k - 1) + (i64)D.n2 * (q - 1), rs, xs, D.r1);
        }
        fa[e] = f;
        fb[e] = res;
        amax_take(braw, f, e);
        amax_take(bres, res, e);
    }
    if (!fold_partials(P, v, braw, bres, shp)) return;
    if (threadIdx.x == 0) {      // dmrgg.f90:527-547, 560-580
        VState& St = P.st[v];
        if (mode == 2) return;
        St.neval += count;
        if (mode == 1) { St.havecol = 1; St.haverow = 1; St.done = 1; return; }
        St.amax = fmax(St.amax, braw.absv);
        if (ISROW) St.haverow = 1; else St.havecol = 1;
        St.crs += 1;
        int done = St.havecol && St.haverow && (St.crs >= 2 * P.piv);
        if (!done) {
            int e = (int)bres.idx;
            if (!ISROW) {
                int j = e / D.r0 + 1, i = e % D.r0 + 1;
                done = St.havecol && St.haverow && (i == St.ii && j == St.jj);
                St.ii = i; St.jj = j;
            } else {
                int q = e / D.n2 + 1, k = e % D.n2 + 1;
                done = St.havecol && St.haverow && (k == St.kk && q == St.qq);
                St.kk = k; St.qq = q;
            }
            St.pivot = bres.val;
        }
        St.done = done;
    }
}

// ----------------------------------------------------------------------------
// K3: full-pivoting superblock (dmrgg.f90:341-396), fused: evaluate a(i,j,k,q), residual against col*row in
// dgemm order (K = r(p)), first-index argmax of |a| and of |b|.  STORE also writes `a` to HBM (HBM-bound variant).
// probe_out != nullptr: measurement entry (fixed bond, results to probe_out instead of the visit state).
// dynamic smem: A[auxsm] | stage
// ----------------------------------------------------------------------------
template <int KIND, int STORE>
__global__ void k_superblock(DevPlan P, int dir, int pp, int fixed_bond, int fixed_v, double* a_out, Partial* probe_out) {
    tl_stamp(P, 2);
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = (fixed_bond > 0) ? fixed_v : P.v0 + blockIdx.y;
    Dims D;
    if (fixed_bond > 0) {
        D.active = 1; D.p = fixed_bond; D.r0 = P.rk[D.p - 1]; D.r1 = P.rk[D.p]; D.r2 = P.rk[D.p + 1];
        D.n1 = P.n[D.p]; D.n2 = P.n[D.p + 1];
    } else {
        if (P.ctrl->ready) return;
        D = load_dims(P, v, dir, pp);
    }
    if (!D.active) return;
    const double* A = stage_aux<KIND>(P, smem);
    double* stg = smem + P.auxsm;
    Stage S;
    if (P.stage) S = stage_bond(P, stg, D.p - 1, D.r0, D.p, D.p + 1, D.p + 1, D.r2);
    const double* colp = P.col + P.coreOff[D.p];
    const double* rowp = P.rowT + P.coreOff[D.p + 1];
    const i64 cs = (i64)P.Rmax * D.n1;
    const i64 rs = (i64)D.n2 * P.Rmax;
    const i64 m1 = (i64)D.r0 * D.n1;           // rows of the unfolding
    const i64 tot = m1 * D.n2 * D.r2;
    Partial braw = amax_init(), bres = amax_init();
    for (i64 x = (i64)blockIdx.x * blockDim.x + threadIdx.x; x < tot; x += (i64)gridDim.x * blockDim.x) {
        i64 kq = x / m1; int ij = (int)(x - kq * m1);
        int q = (int)(kq / D.n2) + 1, k = (int)(kq % D.n2) + 1;
        int j = ij / D.r0 + 1, i = ij % D.r0 + 1;
        double f = eval_bond<KIND>(P, S, D.p, i, j, k, q, A);
        const double* c = colp + (i - 1) + (i64)P.Rmax * (j - 1);
        const double* r = rowp + (k - 1) + (i64)D.n2 * (q - 1);
        double res = f;
        for (int sidx = 0; sidx < D.r1; ++sidx) res = res + (-r[sidx * rs]) * c[sidx * cs];
        if (STORE) a_out[x] = f;
        amax_take(braw, f, x);
        amax_take(bres, res, x);
    }
    if (!fold_partials(P, v, braw, bres, shp)) return;
    if (threadIdx.x == 0) {
        if (probe_out) { probe_out[0] = braw; probe_out[1] = bres; return; }
        VState& St = P.st[v];
        St.amax = fmax(St.amax, braw.absv);
        i64 x = bres.idx;
        i64 kq = x / m1; int ij = (int)(x - kq * m1);
        St.qq = (int)(kq / D.n2) + 1; St.kk = (int)(kq % D.n2) + 1;
        St.jj = ij / D.r0 + 1; St.ii = ij % D.r0 + 1;
        St.pivot = bres.val;
        St.done = 1; St.havecol = 1; St.haverow = 1; St.crs = 0; St.upd = 0;
        St.neval += m1 * D.n2 * D.r2;
    }
}

// ----------------------------------------------------------------------------
// K4: accept test and index-set update (dmrgg.f90:598-660)
// ----------------------------------------------------------------------------
__global__ void k_accept(DevPlan P, int dir, int pp, double small_element, double small_pivot) {
    tl_stamp(P, 3);
    if (P.ctrl->ready) return;
    const int it = P.ctrl->it;
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    VisitOut& O = P.vlog[((i64)(it - 1) * P.maxnb + (pp - 1)) * P.P + v];
    if (!D.active) { if (threadIdx.x == 0) { O.active = 0; O.upd = 0; } return; }
    VState& S = P.st[v];
    __shared__ int s_upd;
    if (threadIdx.x == 0) {
        double ap = fabs(S.pivot);
        int upd = (ap > small_element * S.amax) && (ap > small_pivot * S.pivotmax_prev);
        if (upd && D.r1 >= P.Rmax) { upd = 0; P.ctrl->error = 1; }      // rank capacity (only without maxrank)
        s_upd = upd;
        S.upd = upd;
        O.active = 1; O.upd = upd; O.bond = D.p; O.ii = S.ii; O.jj = S.jj; O.kk = S.kk; O.qq = S.qq; O.pivot = S.pivot;
        if (upd) {
            S.pivotmax = (S.pivotmax < 0.0) ? ap : fmax(S.pivotmax, ap);
            S.pivotmin = (S.pivotmin < 0.0) ? ap : fmin(S.pivotmin, ap);
            int* vp = P.vip + ((i64)D.p * P.Rmax + D.r1) * 4;
            vp[0] = S.ii; vp[1] = S.jj; vp[2] = S.kk; vp[3] = S.qq;
        }
    }
    __syncthreads();
    if (!s_upd) return;
    const int p = D.p, t = D.r1;      // new pivot is number t+1 (0-based column t)
    const int ii = S.ii, jj = S.jj, kk = S.kk, qq = S.qq;
    // flat multi-index tables
    int* Lp = P.Lidx + P.offL[p];
    const int* Lm = P.Lidx + P.offL[p - 1];
    for (int pos = threadIdx.x; pos < p; pos += blockDim.x)
        Lp[(i64)pos * P.Rmax + t] = (pos < p - 1) ? Lm[(i64)pos * P.Rmax + (ii - 1)] : jj;
    int* Rp = P.Ridx + P.offR[p];
    const int* Rn = P.Ridx + P.offR[p + 1];
    for (int pos = threadIdx.x; pos < P.d - p; pos += blockDim.x)
        Rp[(i64)pos * P.Rmax + t] = (pos == 0) ? kk : Rn[(i64)(pos - 1) * P.Rmax + (qq - 1)];
    // packed LU: [ col(ii,jj,1:r) | row(1:r,kk,qq) | pivot ]
    double* g = P.inv + (i64)p * P.Rmax * P.Rmax;
    const double* colp = P.col + P.coreOff[p];
    const double* rowp = P.rowT + P.coreOff[p + 1];
    const i64 cs = (i64)P.Rmax * D.n1, rs = (i64)D.n2 * P.Rmax;
    const int r1 = D.r1;
    for (int s = threadIdx.x; s < r1; s += blockDim.x) {
        g[(i64)r1 * r1 + s] = colp[(ii - 1) + (i64)P.Rmax * (jj - 1) + s * cs];
        g[(i64)r1 * r1 + r1 + s] = rowp[(kk - 1) + (i64)D.n2 * (qq - 1) + s * rs];
    }
    if (threadIdx.x == 0) g[(i64)(r1 + 1) * (r1 + 1) - 1] = S.pivot;
}

// ----------------------------------------------------------------------------
// K5: rank-1 append (dmrgg.f90:663-713).  bcol1/brow1 already hold the lual/luar(from=r+1) eliminations:
// the residual of the last column (row) fiber IS the dgemv of d2_lual (d2_luar) with the same operands and order.
// ----------------------------------------------------------------------------
__global__ void k_update_main(DevPlan P, int dir, int pp) {
    tl_stamp(P, 4);
    if (P.ctrl->ready) return;
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const VState& S = P.st[v];
    if (!S.upd) return;
    const int t = D.r1;   // 0-based slot of the new slice
    double* argp = P.arg + P.coreOff[D.p];
    double* argn = P.arg + P.coreOff[D.p + 1];
    double* colp = P.col + P.coreOff[D.p];
    double* rowp = P.rowT + P.coreOff[D.p + 1];
    const double* acol1 = P.acol1 + (i64)v * P.Rmax * P.nmax;
    const double* bcol1 = P.bcol1 + (i64)v * P.Rmax * P.nmax;
    const double* arow1 = P.arow1 + (i64)v * P.Rmax * P.nmax;
    const double* brow1 = P.brow1 + (i64)v * P.Rmax * P.nmax;
    const double sc = 1.0 / S.pivot;           // dscal(m, 1.d0/g(p**2), ...) (lr.f90:137)
    const int c1 = D.r0 * D.n1, c2 = D.n2 * D.r2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < c1 + c2; e += gridDim.x * blockDim.x) {
        if (e < c1) {
            int j = e / D.r0, i = e % D.r0;
            i64 o = i + (i64)P.Rmax * (j + (i64)D.n1 * t);
            argp[o] = acol1[e];
            colp[o] = sc * bcol1[e];
        } else {
            int x = e - c1;
            int q = x / D.n2, k = x % D.n2;
            argn[t + (i64)P.Rmax * (k + (i64)D.n2 * q)] = arow1[x];
            rowp[k + (i64)D.n2 * (q + (i64)P.Rmax * t)] = brow1[x];
        }
    }
    // r(p) = r(p) + 1 (dmrgg.f90:752) once every CTA of this virtual rank has finished reading the old rank
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tk = atomicAdd(P.tickets + v, 1u);
        s_last = (tk == gridDim.x - 1);
        if (s_last) { P.tickets[v] = 0; P.rk[D.p] = D.r1 + 1; }
    }
}

// neighbour factors (dmrgg.f90:715-749): new column of row(p) through d2_luar(inv(p-1)), new row of col(p+1)
// through d2_lual(inv(p+1)).  One thread per mode index; the triangular recurrences are sequential by definition.
__global__ void k_update_nbr(DevPlan P, int dir, int pp) {
    tl_stamp(P, 5);
    if (P.ctrl->ready) return;
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const VState& S = P.st[v];
    if (!S.upd) return;
    const int lo = P.own[v], hi = P.own[v + 1];
    const int t = D.r1;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < D.n1) {
        if (D.p > lo) {
            // y(1:r0) = arg(p)(:, j, t+1);  y(s) -= sum_{u<s} y(u) * g((s-1)^2 + u)   (dgemv 't' order), stored as rowT(p)(s, j, q=t+1)
            const int j = e;
            const double* acol1 = P.acol1 + (i64)v * P.Rmax * P.nmax;
            const double* g = P.inv + (i64)(D.p - 1) * P.Rmax * P.Rmax;
            double* y = P.rowT + P.coreOff[D.p] + j + (i64)D.n1 * t;    // + s * n1*Rmax
            const i64 ys = (i64)D.n1 * P.Rmax;
            for (int s = 0; s < D.r0; ++s) {
                double val = acol1[s + (i64)D.r0 * j];
                if (s > 0) {
                    double tmp = 0.0;
                    const double* gs = g + (i64)s * s;     // g((s+1-1)^2 + 1 ..), 0-based s
                    for (int u = 0; u < s; ++u) tmp = tmp + y[u * ys] * gs[u];
                    val = val + (-tmp);
                }
                y[s * ys] = val;
            }
        }
    } else if (e - D.n1 < D.n2) {
        if (D.p < hi - 1) {
            // y(1:r2) = arg(p+1)(t+1, k, :);  for c: y(c) += sum_{u<c} (-g(c^2-c+u)) * y(u) (dgemv 'n' order); y(c) *= 1/g(c^2)
            const int k = e - D.n1;
            const double* arow1 = P.arow1 + (i64)v * P.Rmax * P.nmax;
            const double* g = P.inv + (i64)(D.p + 1) * P.Rmax * P.Rmax;
            double* y = P.col + P.coreOff[D.p + 1] + t + (i64)P.Rmax * k;   // + c * Rmax*n2
            const i64 ys = (i64)P.Rmax * D.n2;
            for (int c = 0; c < D.r2; ++c) {
                double val = arow1[k + (i64)D.n2 * c];
                const double* gc = g + (i64)(c + 1) * (c + 1) - (c + 1);   // g(c1^2 - c1 + 1 ..) with c1 = c+1
                for (int u = 0; u < c; ++u) val = val + (-gc[u]) * y[u * ys];
                val = (1.0 / g[(i64)(c + 1) * (c + 1) - 1]) * val;
                y[c * ys] = val;
            }
        }
    }
}

// ----------------------------------------------------------------------------
// sweep begin / end
// ----------------------------------------------------------------------------
// MPI_ALLREDUCE(MAX) of (amax, pivotmax, -pivotmin) (dmrgg.f90:852-870); single thread, P is small
__global__ void k_allreduce(DevPlan P) {
    tl_stamp(P, 6);
    if (P.ctrl->ready) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (P.P > 1) {
        double c1 = P.st[0].amax, c2 = P.st[0].pivotmax, c3 = (P.st[0].pivotmin > 0.0) ? -P.st[0].pivotmin : -999e9;
        for (int v = 1; v < P.P; ++v) {
            c1 = fmax(c1, P.st[v].amax); c2 = fmax(c2, P.st[v].pivotmax);
            c3 = fmax(c3, (P.st[v].pivotmin > 0.0) ? -P.st[v].pivotmin : -999e9);
        }
        for (int v = 0; v < P.P; ++v) {
            P.st[v].amax = c1; P.st[v].pivotmax = c2; P.st[v].pivotmin = -c3;
            if (P.st[v].pivotmin == 999e9) P.st[v].pivotmin = -1.0;
        }
    }
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_run_begin(DevPlan P, unsigned long long seed, int has_accuracy, double accuracy, unsigned long long run_serial) {
    tl_stamp(P, 7);
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    P.ctrl->run_serial = run_serial;
    P.ctrl->ready = 0; P.ctrl->strike = 0; P.ctrl->error = 0; P.ctrl->nsweeps = 0; P.ctrl->it = 1; P.ctrl->itq = 0;
    P.ctrl->seed = seed; P.ctrl->has_accuracy = has_accuracy; P.ctrl->accuracy = accuracy;
    P.ctrl->t0_ns = globaltimer_ns();
    for (int v = 0; v <= P.P; ++v) P.tickets[v] = 0;
    for (int x = 0; x < 2 * (P.d + 1); ++x) P.qext[x] = 0;
}
// end of sweep `it`: the scalar reductions of dmrgg.f90:961-967, the record of the sweep (after the quadrature), the exit
// test of dmrgg.f90:1010-1019, and the preparation of the next sweep (rr = r snapshot of :325, pivotmax = pivotmin = -1)
// (block-cooperative body: every thread of ONE CTA calls it; also run by the last quadrature kernel of a sweep)
__device__ __forceinline__ void sweep_log_body(const DevPlan& P, int maxrank, bool with_val = true) {
    const int it = P.ctrl->it;
    __shared__ unsigned long long s_ne;
    __shared__ double s_amax, s_pmax, s_pmin;
    if (threadIdx.x == 0) { s_ne = 0ULL; s_amax = P.st[0].amax; s_pmax = P.st[0].pivotmax; s_pmin = P.st[0].pivotmin; }
    for (int x = threadIdx.x; x <= P.d; x += blockDim.x) {
        int r = P.rk[x];
        P.rklog[(i64)it * (P.d + 1) + x] = r; P.rks[x] = r; P.qsnap[(it & 1) * (P.d + 1) + x] = r;
    }
    __syncthreads();
    for (int v = threadIdx.x; v < P.P; v += blockDim.x) {
        atomicAdd(&s_ne, (unsigned long long)P.st[v].neval);
        P.st[v].pivotmax_prev = P.st[v].pivotmax;        // dmrgg.f90:961
        P.st[v].pivotmax = -1.0; P.st[v].pivotmin = -1.0; // dmrgg.f90:326-327 of the next sweep
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const i64 ne = (i64)s_ne;
    SweepOut o;
    o.neval = ne; o.amax = s_amax; o.pivotmax = s_pmax; o.pivotmin = s_pmin;
    o.t_ns = globaltimer_ns() - P.ctrl->t0_ns;
    if (with_val) P.slog[it].val = P.sweep_out->val;     // (overlapped quadrature: quad_record() fills val later)
    P.slog[it].neval = o.neval; P.slog[it].amax = o.amax; P.slog[it].pivotmax = o.pivotmax; P.slog[it].pivotmin = o.pivotmin;
    P.slog[it].t_ns = o.t_ns; P.slog[it].valid = 1; P.slog[it].pad = 0;
    P.ctrl->nsweeps = it;
    int ready = 0;
    if (maxrank > 0) ready = (it + 1 >= maxrank);
    if (P.ctrl->has_accuracy) {
        if (o.pivotmax <= P.ctrl->accuracy * o.amax) P.ctrl->strike += 1; else P.ctrl->strike = 0;
        ready = ready || (P.ctrl->strike >= 3);
    }
    if (P.ctrl->error) ready = 1;
    P.ctrl->it = it + 1;
    __threadfence();
    P.ctrl->ready = ready;
}
__global__ void k_sweep_log(DevPlan P, int maxrank, int with_val) {
    tl_stamp(P, 8);
    if (P.ctrl->ready) return;
    sweep_log_body(P, maxrank, with_val != 0);
}
// Overlapped quadrature (log_maxrank < 0 in the quadrature kernels): the exit test of dmrgg.f90:1010-1019 does not use the
// quadrature value, so k_sweep_log(with_val = 0) closes sweep s right after the exchange and the quadrature of sweep s
// runs on a second stream beside the bond visits of sweep s+1, reading the rank snapshot qsnap[s & 1].  ctrl->it - 1
// sweeps are closed, ctrl->itq of them have their value recorded (itq changes only at the end of a quadrature group; the
// close of sweep s+2 waits for the quadrature of sweep s by an event, so its snapshot buffer is not overwritten early).
__device__ __forceinline__ bool quad_pending(const DevPlan& P) { return P.ctrl->itq < *(volatile int*)&P.ctrl->it - 1; }
__device__ __forceinline__ const int* quad_ranks(const DevPlan& P, int ovl) {
    return ovl ? P.qsnap + ((P.ctrl->itq + 1) & 1) * (P.d + 1) : P.rk;
}
__device__ __forceinline__ void quad_record(const DevPlan& P, double val) {    // one thread of the group's last kernel
    if (!quad_pending(P)) return;
    const int sw = P.ctrl->itq + 1;
    P.slog[sw].val = val;
    __threadfence();
    P.ctrl->itq = sw;
}

// ----------------------------------------------------------------------------
// neighbour exchange between virtual ranks (dmrgg.f90:872-958 LEFT, dmrggmp.f90:572-629 RIGHT).
// With one padded copy of every core in HBM the new column / row of the shared core is already in place;
// what remains is the corner (both sides evaluate the same n(c) points; each counts them) and the factor extensions.
// blockIdx.y = boundary b between virtual ranks b and b+1, shared core c = own[b+1].
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_exchange_corner(DevPlan P) {
    tl_stamp(P, 9);
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int b = first_boundary(P) + blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    if (!(rc1 > rc1s && rc > rcs)) return;
    const double* A = stage_aux<KIND>(P, smem);
    const int nc = P.n[c];
    double* argc = P.arg + P.coreOff[c];
    // the d-1 fixed positions of the corner fiber, staged once (one round trip instead of a pointer chase per evaluation)
    double* XF = smem + P.auxsm; double* WF = XF + P.d;
    {
        const int* Lt = P.Lidx + P.offL[c - 1]; const int* Rt = P.Ridx + P.offR[c];
        const bool hasw = (P.kind == KIND_ISING);
        const int nwoff = P.n[1];
        for (int pos = threadIdx.x; pos < P.d - 1; pos += blockDim.x) {
            const int idx = (pos < c - 1) ? Lt[(i64)pos * P.Rmax + (rc1 - 1)] : Rt[(i64)(pos - (c - 1)) * P.Rmax + (rc - 1)];
            XF[pos] = P.par[idx - 1]; WF[pos] = hasw ? P.par[nwoff + idx - 1] : 0.0;
        }
        __syncthreads();
    }
    Partial best = amax_init();
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        StagedVals sv;
        sv.XL = XF; sv.WL = WF; sv.nl = c - 1; sv.rl = 1; sv.i = 1;
        sv.xj = P.par[j]; sv.wj = (P.kind == KIND_ISING) ? P.par[P.n[1] + j] : 0.0;
        sv.hask = 0; sv.xk = 0.0; sv.wk = 0.0;
        sv.XR = XF + (c - 1); sv.WR = WF + (c - 1); sv.rr = 1; sv.q = 1;
        double f = eval_point_wide<KIND>(P, sv, A);
        argc[(rc1 - 1) + (i64)P.Rmax * (j + (i64)nc * (rc - 1))] = f;
        amax_take(best, f, j);
    }
    best = amax_block(best, shp);
    if (threadIdx.x == 0) {
        // virtual rank b+1 is also touched by the CTA of boundary b+1: atomics (amax >= 0, so the bit pattern orders like the value)
        for (int v = b; v <= b + 1; ++v) {
            if (!own_vrank(P, v)) continue;     // the other side of a process boundary evaluates (and counts) its own copy
            atomicMax((long long*)&P.st[v].amax, __double_as_longlong(best.absv));
            atomicAdd((unsigned long long*)&P.st[v].neval, (unsigned long long)nc);
        }
    }
}
__global__ void k_exchange_extend(DevPlan P) {
    tl_stamp(P, 10);
    if (P.ctrl->ready) return;
    const int b = first_boundary(P) + blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    const int nc = P.n[c];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const double* argc = P.arg + P.coreOff[c];
    if (e < nc) {
        if (rc > rcs && own_vrank(P, b)) {
            // LEFT receiver (virtual rank b): row(c)(:, k, rc) = d2_luar(n(c), rc1, inv(c-1)) of arg(c)(:, :, rc)
            const int k = e;
            const double* g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax;
            double* y = P.rowT + P.coreOff[c] + k + (i64)nc * (rc - 1);
            const i64 ys = (i64)nc * P.Rmax;
            const double* a = argc + (i64)P.Rmax * (k + (i64)nc * (rc - 1));
            for (int s = 0; s < rc1; ++s) {
                double val = a[s];
                if (s > 0) {
                    double tmp = 0.0;
                    const double* gs = g + (i64)s * s;
                    for (int u = 0; u < s; ++u) tmp = tmp + y[u * ys] * gs[u];
                    val = val + (-tmp);
                }
                y[s * ys] = val;
            }
        }
    } else if (e - nc < nc) {
        if (rc1 > rc1s && own_vrank(P, b + 1)) {
            // RIGHT receiver (virtual rank b+1): col(c)(rc1, j, :) = d2_lual(n(c), rc, inv(c)) of arg(c)(rc1, :, :)
            const int j = e - nc;
            const double* g = P.inv + (i64)c * P.Rmax * P.Rmax;
            double* y = P.col + P.coreOff[c] + (rc1 - 1) + (i64)P.Rmax * j;
            const i64 ys = (i64)P.Rmax * nc;
            const double* a = argc + (rc1 - 1) + (i64)P.Rmax * j;
            for (int cc = 0; cc < rc; ++cc) {
                double val = a[cc * ys];
                const double* gc = g + (i64)(cc + 1) * (cc + 1) - (cc + 1);
                for (int u = 0; u < cc; ++u) val = val + (-gc[u]) * y[u * ys];
                val = (1.0 / g[(i64)(cc + 1) * (cc + 1) - 1]) * val;
                y[cc * ys] = val;
            }
        }
    }
}

// ----------------------------------------------------------------------------
// quadrature (dmrgg.f90:975-993 per sweep; dtt_lua :1169-1258; dtt_quad :1261-1415)
// ----------------------------------------------------------------------------
// ttqq(p)(i,k) = sum_j arg(p)(i,j,k) * w_p(j), accumulated from 0 in ascending j (dgemv 'n', beta = 0)
__global__ void k_quad_contract(DevPlan P, int use_weights) {
    tl_stamp(P, 11);
    const int p = P.c_lo + blockIdx.y;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const double* a = P.arg + P.coreOff[p];
    const double* w = P.quadw + P.quadOff[p];
    double* out = P.ttqq + (i64)p * P.Rmax * P.Rmax;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < r0 * r1; e += gridDim.x * blockDim.x) {
        int k = e / r0, i = e % r0;
        const double* c = a + i + (i64)P.Rmax * n * k;
        double y = 0.0;
        if (use_weights) for (int j = 0; j < n; ++j) y = y + w[j] * c[(i64)P.Rmax * j];
        else             for (int j = 0; j < n; ++j) y = y + c[(i64)P.Rmax * j];
        out[i + (i64)P.Rmax * k] = y;
    }
}
// dtt_lua on the contracted train: core p is an r0 x r1 matrix with leading dimension Rmax.
// One CTA per core: d2_luar over columns (thread per column), then d2_lual over rows (thread per row).
__global__ void k_quad_lua(DevPlan P) {
    tl_stamp(P, 12);
    if (P.ctrl->ready) return;
    const int p = P.c_lo + blockIdx.x;
    const int r0 = P.rk[p - 1], r1 = P.rk[p];
    double* m = P.ttqq + (i64)p * P.Rmax * P.Rmax;
    const double* gl = P.inv + (i64)(p - 1) * P.Rmax * P.Rmax;
    const double* gr = P.inv + (i64)p * P.Rmax * P.Rmax;
    for (int k = threadIdx.x; k < r1; k += blockDim.x) {
        double* y = m + (i64)P.Rmax * k;
        for (int s = 1; s < r0; ++s) {
            double tmp = 0.0;
            const double* gs = gl + (i64)s * s;
            for (int u = 0; u < s; ++u) tmp = tmp + y[u] * gs[u];
            y[s] = y[s] + (-tmp);
        }
    }
    __syncthreads();
    if (p < P.d) {
        for (int i = threadIdx.x; i < r0; i += blockDim.x) {
            double* y = m + i;
            for (int c = 0; c < r1; ++c) {
                double val = y[(i64)P.Rmax * c];
                const double* gc = gr + (i64)(c + 1) * (c + 1) - (c + 1);
                for (int u = 0; u < c; ++u) val = val + (-gc[u]) * y[(i64)P.Rmax * u];
                val = (1.0 / gr[(i64)(c + 1) * (c + 1) - 1]) * val;
                y[(i64)P.Rmax * c] = val;
            }
        }
    }
}
// chain product of each virtual rank's cores (dmrgg.f90:1323-1345), one CTA per virtual rank, then the binary tree
// over virtual ranks (:1355-1405) by CTA 0 of a second launch.  All matrices have leading dimension Rmax.
__device__ __forceinline__ void mat_mul(const double* A, int m, int kdim, const double* B, int n, double* C, int ld) {
    // C(m x n) = A(m x kdim) * B(kdim x n), dgemm 'n','n' order: c(i,j) += b(l,j) * a(i,l), l ascending, from 0
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
        int j = e / m, i = e % m;
        double c = 0.0;
        for (int l = 0; l < kdim; ++l) c = c + B[l + (i64)ld * j] * A[i + (i64)ld * l];
        C[i + (i64)ld * j] = c;
    }
}
__global__ void k_quad_chain(DevPlan P) {
    tl_stamp(P, 13);
    const int v = P.v0 + blockIdx.x;
    const int first = P.own[v];
    int last = P.own[v + 1] - 1;
    if (v == P.P - 1) last = P.d;
    const i64 msz = (i64)P.Rmax * P.Rmax;
    double* cur = P.chain + (i64)v * msz;
    double* tmp = P.chain2 + (i64)v * msz;
    const int m = P.rk[first - 1];
    {
        const double* src = P.ttqq + (i64)first * msz;
        for (int e = threadIdx.x; e < msz; e += blockDim.x) cur[e] = src[e];
    }
    __syncthreads();
    for (int p = first + 1; p <= last; ++p) {
        mat_mul(cur, m, P.rk[p - 1], P.ttqq + (i64)p * msz, P.rk[p], tmp, P.Rmax);
        __syncthreads();
        for (int e = threadIdx.x; e < msz; e += blockDim.x) cur[e] = tmp[e];
        __syncthreads();
    }
}
__global__ void k_quad_tree(DevPlan P) {
    tl_stamp(P, 14);
    const i64 msz = (i64)P.Rmax * P.Rmax;
    for (int q = 1; q < P.P; q *= 2) {
        for (int me = 0; me < P.P; me += 2 * q) {
            int her = me + q;
            if (her < P.P) {
                double* a = P.chain + (i64)me * msz;
                const double* b = P.chain + (i64)her * msz;
                double* tmp = P.chain2 + (i64)me * msz;
                // dims: a is r(own[me]-1) x r(own[her]-1) ; b is r(own[her]-1) x r(end of her's accumulated span)
                int herend = her + q; if (herend > P.P) herend = P.P;
                int m = P.rk[P.own[me] - 1], kd = P.rk[P.own[her] - 1];
                int n = (herend == P.P) ? P.rk[P.d] : P.rk[P.own[herend] - 1];
                mat_mul(a, m, kd, b, n, tmp, P.Rmax);
                __syncthreads();
                for (int e = threadIdx.x; e < msz; e += blockDim.x) a[e] = tmp[e];
                __syncthreads();
            }
        }
    }
    if (threadIdx.x == 0) P.sweep_out->val = P.chain[0];
}

// finalisation: dtt_lua on the real cores, in place (dmrgg.f90:1248-1257)
__global__ void k_lua_r(DevPlan P) {
    tl_stamp(P, 15);   // d2_luar(n*r1, r0, inv(p-1)): thread per column (j,k)
    const int p = P.c_lo + blockIdx.y;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    if (r0 < 2) return;
    const double* g = P.inv + (i64)(p - 1) * P.Rmax * P.Rmax;
    double* a = P.arg + P.coreOff[p];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * r1; e += gridDim.x * blockDim.x) {
        int k = e / n, j = e % n;
        double* y = a + (i64)P.Rmax * (j + (i64)n * k);
        for (int s = 1; s < r0; ++s) {
            double tmp = 0.0;
            const double* gs = g + (i64)s * s;
            for (int u = 0; u < s; ++u) tmp = tmp + y[u] * gs[u];
            y[s] = y[s] + (-tmp);
        }
    }
}
__global__ void k_lua_l(DevPlan P) {
    tl_stamp(P, 16);   // d2_lual(r0*n, r1, inv(p)): thread per row (i,j); cores 1..d-1
    const int p = P.c_lo + blockIdx.y;
    if (p >= P.d) return;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const double* g = P.inv + (i64)p * P.Rmax * P.Rmax;
    double* a = P.arg + P.coreOff[p];
    const i64 ys = (i64)P.Rmax * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < r0 * n; e += gridDim.x * blockDim.x) {
        int j = e / r0, i = e % r0;
        double* y = a + i + (i64)P.Rmax * j;
        for (int c = 0; c < r1; ++c) {
            double val = y[c * ys];
            const double* gc = g + (i64)(c + 1) * (c + 1) - (c + 1);
            for (int u = 0; u < c; ++u) val = val + (-gc[u]) * y[u * ys];
            val = (1.0 / g[(i64)(c + 1) * (c + 1) - 1]) * val;
            y[c * ys] = val;
        }
    }
}
// padded -> packed copy of one core for ttc_core()
__global__ void k_pack_core(DevPlan P, int p, double* out) {
    tl_stamp(P, 17);
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const double* a = P.arg + P.coreOff[p];
    const i64 tot = (i64)r0 * n * r1;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        int i = (int)(e % r0); i64 jk = e / r0;
        out[e] = a[i + (i64)P.Rmax * jk];
    }
}
// padded -> packed copy of EVERY core c_lo..c_hi this process owns, concatenated in core order (the layout of ttc_cores), with the
// offsets taken from the device-side ranks: enqueued right behind the finalisation, before the host knows the ranks
// (ttc_bind_cores).  grid (x, number of own cores).
__global__ void k_pack_all(DevPlan P, double* out) {
    const int p = P.c_lo + blockIdx.y;
    i64 off = 0;
    for (int k = P.c_lo; k < p; ++k) off += (i64)P.rk[k - 1] * P.n[k] * P.rk[k];
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const double* a = P.arg + P.coreOff[p];
    const i64 tot = (i64)r0 * n * r1;
    double* o = out + off;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        int i = (int)(e % r0); i64 jk = e / r0;
        o[e] = a[i + (i64)P.Rmax * jk];
    }
}

// ----------------------------------------------------------------------------
// initial cross (dmrgg.f90:150-232)
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_init_search(DevPlan P, int nn, int snum, double* b) {
    tl_stamp(P, 18);
    extern __shared__ double smem[];
    const double* A = stage_aux<KIND>(P, smem);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nn * snum; x += gridDim.x * blockDim.x) {
        DiagSrc s; s.n = P.n; s.k = x % nn + 1; s.s = x / nn;
        b[x] = eval_src<KIND>(P, s, A);
    }
}
// fiber of core p through the initial cross: arg(p)(1,j,1) = f(ind0 with position p := j); tables hold pivot 1 already
template <int KIND>
__global__ void k_init_cross(DevPlan P) {
    tl_stamp(P, 19);
    extern __shared__ double smem[];
    const int p = blockIdx.y + 1;
    const double* A = stage_aux<KIND>(P, smem);
    const int n = P.n[p];
    double* a = P.arg + P.coreOff[p];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        PointSrc s;
        s.L = P.Lidx + P.offL[p - 1]; s.nl = p - 1; s.i = 1; s.j = j + 1; s.k = 0; s.hask = 0;
        s.R = P.Ridx + P.offR[p]; s.q = 1; s.Rmax = P.Rmax;
        a[(i64)P.Rmax * j] = eval_src<KIND>(P, s, A);
    }
}

// =============================================================================
// Warp-cooperative triangular recurrences (d2_luar / d2_lual, lr.f90:124-154).
//
// The recurrences are sequential by definition, but only along ONE chain per output: row s of d2_luar needs
// tmp_s = sum_{u<s} y(u)*g(s,u) accumulated in ascending u, and y(u) is final after step u-1.  A warp therefore runs
// the chain as a wavefront: lane s keeps tmp_s, at step u the final y(u) is broadcast and every lane s > u adds its
// term — the same additions in the same order as the reference, r steps instead of r^2/2, no memory on the chain.
// Lanes own rows s = lane + 32*t, t < MAXRPL (r <= 32*MAXRPL).
// =============================================================================
constexpr int MAXRPL = 4;
constexpr unsigned FULLMASK = 0xffffffffu;

template <class GF>
__device__ __forceinline__ void warp_luar(double (&y)[MAXRPL], int r, GF g) {
    const int lane = threadIdx.x & 31;
    if (r <= 32) {          // one row per lane: no slot selection on the chain, coefficient loads issued ahead of it
        double y0 = y[0], tmp0 = 0.0;
        const bool in = lane < r;
#pragma unroll 4
        for (int u = 0; u + 1 < r; ++u) {
            const double gsu = (in && lane > u) ? g(lane, u) : 0.0;
            const double yu = __shfl_sync(FULLMASK, y0, u);
            if (in && lane > u) tmp0 = tmp0 + yu * gsu;
            if (lane == u + 1) y0 = y0 + (-tmp0);
        }
        y[0] = y0;
        return;
    }
    double tmp[MAXRPL];
#pragma unroll
    for (int t = 0; t < MAXRPL; ++t) tmp[t] = 0.0;
    for (int u = 0; u + 1 < r; ++u) {
        double yu = 0.0;
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) if (t == (u >> 5)) yu = y[t];
        yu = __shfl_sync(FULLMASK, yu, u & 31);
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) {
            const int s = lane + 32 * t;
            if (s > u && s < r) tmp[t] = tmp[t] + yu * g(s, u);
        }
        const int s1 = u + 1;
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) if (t == (s1 >> 5) && lane == (s1 & 31)) y[t] = y[t] + (-tmp[t]);
    }
}
// d2_lual on one row: val(c) = y(c) + sum_{u<c} (-g(c,u))*y(u) (ascending u), then val(c) *= dinv(c); gl(c,u), dinv(c) functors
template <class GF, class DF>
__device__ __forceinline__ void warp_lual(double (&y)[MAXRPL], int r, GF g, DF dinv) {
    const int lane = threadIdx.x & 31;
    if (r <= 32) {
        double y0 = y[0];
        const bool in = lane < r;
        const double di = in ? dinv(lane) : 0.0;
        if (r > 0 && lane == 0) y0 = di * y0;
#pragma unroll 4
        for (int u = 0; u + 1 < r; ++u) {
            const double gcu = (in && lane > u) ? g(lane, u) : 0.0;
            const double yu = __shfl_sync(FULLMASK, y0, u);
            if (in && lane > u) y0 = y0 + (-gcu) * yu;
            if (lane == u + 1) y0 = di * y0;
        }
        y[0] = y0;
        return;
    }
    if (r > 0 && lane == 0) y[0] = dinv(0) * y[0];
    for (int u = 0; u + 1 < r; ++u) {
        double yu = 0.0;
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) if (t == (u >> 5)) yu = y[t];
        yu = __shfl_sync(FULLMASK, yu, u & 31);
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) {
            const int c = lane + 32 * t;
            if (c > u && c < r) y[t] = y[t] + (-g(c, u)) * yu;
        }
        const int c1 = u + 1;
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) if (t == (c1 >> 5) && lane == (c1 & 31)) y[t] = dinv(c1) * y[t];
    }
}
// packed LU accessors (dmrgg.f90:650-660): row part g(s^2 + u), column part g((c+1)^2 - (c+1) + u), pivot g((c+1)^2 - 1)
struct GLuar { const double* g; __device__ __forceinline__ double operator()(int s, int u) const { return g[(i64)s * s + u]; } };
struct GLual { const double* g; __device__ __forceinline__ double operator()(int c, int u) const { return g[(i64)(c + 1) * (c + 1) - (c + 1) + u]; } };
struct GDinv { const double* g; __device__ __forceinline__ double operator()(int c) const { return 1.0 / g[(i64)(c + 1) * (c + 1) - 1]; } };
// transposed shared-memory copies: T[u*ld + s]; lanes read consecutive s -> conflict free
struct GSm { const double* T; int ld; __device__ __forceinline__ double operator()(int s, int u) const { return T[u * ld + s]; } };
struct DSm { const double* d; __device__ __forceinline__ double operator()(int c) const { return d[c]; } };
__device__ __forceinline__ void stage_luar(const double* g, int r, double* T) {
    for (int x = threadIdx.x; x < r * r; x += blockDim.x) { int s = x / r, u = x - s * r; if (u < s) T[u * r + s] = g[(i64)s * s + u]; }
}
__device__ __forceinline__ void stage_lual(const double* g, int r, double* T, double* dinv) {
    for (int x = threadIdx.x; x < r * r; x += blockDim.x) { int c = x / r, u = x - c * r; if (u < c) T[u * r + c] = g[(i64)(c + 1) * (c + 1) - (c + 1) + u]; }
    for (int c = threadIdx.x; c < r; c += blockDim.x) dinv[c] = 1.0 / g[(i64)(c + 1) * (c + 1) - 1];
}

// ----------------------------------------------------------------------------
// quadrature, shared-memory versions (same arithmetic as k_quad_contract / k_quad_lua / k_quad_chain / k_quad_tree)
// ----------------------------------------------------------------------------
// ttqq(p)(i,k) = sum_j arg(p)(i,j,k)*w(j), j ascending (dgemv order, dmrgg.f90:986-991).  A warp owns one k, a lane one i: the
// left index is contiguous in HBM, so every step of the ordered sum is one coalesced 256 B row; the rows are fetched QC_U at a
// time into registers, the next batch in flight while the current one is summed, so the dependent additions never wait for
// L2.  Two warps per CTA spread the 13 MB of config B over all SMs.  (The first version staged a whole slice in shared memory
// with 256 threads and let r0 of them do the sums: 28 us per launch at config B, two launches per run.)
constexpr int QC_U = 32, QC_WARPS = 2;
__global__ void __launch_bounds__(32 * QC_WARPS) k_quad_contract_sm(DevPlan P, int use_weights) {
    tl_stamp(P, 20);
    extern __shared__ double smem[];           // the weights of this core (nmax doubles)
    const int p = P.c_lo + blockIdx.y;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    if (use_weights) { const double* w = P.quadw + P.quadOff[p]; for (int j = threadIdx.x; j < n; j += blockDim.x) smem[j] = w[j]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, k = blockIdx.x * QC_WARPS + (threadIdx.x >> 5);
    if (k >= r1) return;
    const i64 ld = P.Rmax;
    for (int i = lane; i < r0; i += 32) {
        const double* a = P.arg + P.coreOff[p] + ld * n * k + i;
        double y = 0.0, cur[QC_U], nxt[QC_U];
        const int nb = n / QC_U;
        if (nb > 0) {
#pragma unroll
            for (int u = 0; u < QC_U; ++u) cur[u] = a[ld * u];
        }
        for (int b = 0; b < nb; ++b) {
            const int j0 = b * QC_U;
            if (b + 1 < nb) {
#pragma unroll
                for (int u = 0; u < QC_U; ++u) nxt[u] = a[ld * (j0 + QC_U + u)];
            }
            if (use_weights) {
#pragma unroll
                for (int u = 0; u < QC_U; ++u) y = y + smem[j0 + u] * cur[u];
            } else {
#pragma unroll
                for (int u = 0; u < QC_U; ++u) y = y + cur[u];
            }
#pragma unroll
            for (int u = 0; u < QC_U; ++u) cur[u] = nxt[u];
        }
        for (int j = nb * QC_U; j < n; ++j) y = use_weights ? y + smem[j] * a[ld * j] : y + a[ld * j];
        P.ttqq[(i64)p * P.Rmax * P.Rmax + i + ld * k] = y;
    }
}
// dtt_lua on the contracted cores: one CTA per core, matrix and both packed LUs staged in shared memory,
// one warp per column (d2_luar) then one warp per row (d2_lual).  Requires r <= 32*MAXRPL.
__global__ void k_quad_lua_sm(DevPlan P) {
    tl_stamp(P, 21);
    extern __shared__ double smem[];
    if (P.ctrl->ready) return;
    const int p = P.c_lo + blockIdx.x;
    const int r0 = P.rk[p - 1], r1 = P.rk[p];
    double* M = smem;                          // r0 x r1, ld r0
    double* TL = M + r0 * r1;                  // luar table of inv(p-1), r0 x r0
    double* TR = TL + r0 * r0;                 // lual table of inv(p), r1 x r1
    double* DI = TR + r1 * r1;                 // r1
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
// C(m x n) = A(m x kd) * B(kd x n), all in shared memory with leading dimension ld, dgemm order
__device__ __forceinline__ void mat_mul_sm(const double* A, int m, int kd, const double* B, int n, double* C, int ld) {
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) {
        int j = e / m, i = e - j * m;
        double c = 0.0;
        for (int l = 0; l < kd; ++l) c = c + B[l + ld * j] * A[i + ld * l];
        C[i + ld * j] = c;
    }
}
__device__ __forceinline__ void mat_load_sm(const double* g, int m, int n, int ldg, double* S, int ld) {
    for (int e = threadIdx.x; e < m * n; e += blockDim.x) { int j = e / m, i = e - j * m; S[i + ld * j] = g[i + (i64)ldg * j]; }
}
// chain product per virtual rank (dmrgg.f90:1323-1345): CTA v, three shared buffers of Rmax^2
__global__ void k_quad_chain_sm(DevPlan P, int log_maxrank) {
    tl_stamp(P, 22);
    extern __shared__ double smem[];
    const int v = P.v0 + blockIdx.x;
    const int first = P.own[v];
    int last = P.own[v + 1] - 1;
    if (v == P.P - 1) last = P.d;
    const int ld = P.Rmax;
    const int* rq = quad_ranks(P, log_maxrank < 0);
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
    double* out = P.chain + (i64)v * msz;
    const int nl = rq[last];
    for (int e = threadIdx.x; e < m * nl; e += blockDim.x) { int j = e / m, i = e - j * m; out[i + (i64)ld * j] = cur[i + ld * j]; }
    if (P.P == 1 && threadIdx.x == 0) { P.sweep_out->val = cur[0]; if (log_maxrank < 0) quad_record(P, cur[0]); }
    if (P.P == 1 && log_maxrank > 0 && !P.ctrl->ready) { __syncthreads(); sweep_log_body(P, log_maxrank); }   // single partition: this is the sweep's last kernel
}
// binary tree over virtual ranks (dmrgg.f90:1355-1405): level `q`, CTA per receiving rank; launched once per level
__global__ void k_quad_tree_sm(DevPlan P, int q, int last_level, int log_maxrank) {
    tl_stamp(P, 23);
    extern __shared__ double smem[];
    const int me = blockIdx.x * 2 * q, her = me + q;
    const int ld = P.Rmax;
    const int* rq = quad_ranks(P, log_maxrank < 0);
    const i64 msz = (i64)ld * ld;
    if (her < P.P) {
        double* A = smem; double* B = smem + msz; double* C = smem + 2 * msz;
        int herend = her + q; if (herend > P.P) herend = P.P;
        const int m = rq[P.own[me] - 1], kd = rq[P.own[her] - 1];
        const int n = (herend == P.P) ? rq[P.d] : rq[P.own[herend] - 1];
        mat_load_sm(P.chain + (i64)me * msz, m, kd, ld, A, ld);
        mat_load_sm(P.chain + (i64)her * msz, kd, n, ld, B, ld);
        __syncthreads();
        mat_mul_sm(A, m, kd, B, n, C, ld);
        __syncthreads();
        double* out = P.chain + (i64)me * msz;
        for (int e = threadIdx.x; e < m * n; e += blockDim.x) { int j = e / m, i = e - j * m; out[i + (i64)ld * j] = C[i + ld * j]; }
        if (last_level && me == 0 && threadIdx.x == 0) { P.sweep_out->val = C[0]; if (log_maxrank < 0) quad_record(P, C[0]); }
    }
    // the root of the tree is the sweep's last kernel: it also writes the sweep record and takes the exit decision
    if (last_level && me == 0 && log_maxrank > 0 && !P.ctrl->ready) { __syncthreads(); sweep_log_body(P, log_maxrank); }
}

// ----------------------------------------------------------------------------
// Incremental per-sweep quadrature (dmrgg.f90:975-993: ttqq = arg x weights, dtt_lua(ttqq), dtt_quad(ttqq)).
// A sweep appends at most one row and one column to every contracted core; every entry of the contracted,
// luar'd and lual'd core is a fixed sequence of operations on data that never changes once written (the raw
// fibers and the packed LUs only grow), so the old entries are bit-identical to a recomputation and only the
// new row i = e0 and the new column k = e1 are evaluated: O((r0 + r1) * (n + r)) instead of O(r0 * r1 * (n + r)).
//   Y = after d2_luar(inv(p-1)) on every column, Z = after d2_lual(inv(p)) on every row (Z = Y for the last core).
// One CTA per own core; qext[p] = extents already done.  Needs r <= 32*MAXRPL.
// ----------------------------------------------------------------------------
__global__ void k_quad_inc(DevPlan P, int use_weights, int stage_doubles, int ovl) {
    tl_stamp(P, 34);
    if (ovl ? !quad_pending(P) : (P.ctrl->ready != 0)) return;
    extern __shared__ double smem[];
    const int p = P.c_lo + blockIdx.x;
    const int* rq = quad_ranks(P, ovl);
    const int r0 = rq[p - 1], r1 = rq[p], n = P.n[p];
    const int e0 = P.qext[2 * p], e1 = P.qext[2 * p + 1];
    if (e0 == r0 && e1 == r1) return;
    if (r0 - e0 > 1 || r1 - e1 > 1 || r0 < e0 || r1 < e1) { if (threadIdx.x == 0) P.ctrl->error = 2; return; }
    const bool grow_row = r0 > e0, grow_col = r1 > e1;
    const double* a = P.arg + P.coreOff[p];
    const double* w = P.quadw + P.quadOff[p];
    double* Y = P.ttqy + (i64)p * P.Rmax * P.Rmax;
    double* Z = P.ttqq + (i64)p * P.Rmax * P.Rmax;
    const double* gl = P.inv + (i64)(p - 1) * P.Rmax * P.Rmax;
    const double* gr = P.inv + (i64)p * P.Rmax * P.Rmax;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int ld = P.Rmax, R = P.Rmax;
    // shared: TL[R*R] luar table of inv(p-1) | TR[R*R] lual table of inv(p) | DI[R] | rawcol[R] | rawrow[R] |
    //         YS[R*R] old Y (ld e0) | ZS[R*R] old Z (ld e0) | stage[...]
    double* TL = smem; double* TR = TL + R * R; double* DI = TR + R * R; double* rawcol = DI + R; double* rawrow = rawcol + R;
    double* YS = rawrow + R; double* ZS = YS + R * R; double* stg = ZS + R * R;
    stage_luar(gl, r0, TL);
    if (p < P.d) stage_lual(gr, r1, TR, DI);
    for (int k = wid; k < e1; k += nw)
        for (int u = lane; u < e0; u += 32) { YS[u + e0 * k] = Y[u + ld * k]; ZS[u + e0 * k] = Z[u + ld * k]; }
    tl_mark0(P, 60);
    // 1. raw contractions of the new column (all rows) and of the new row (all columns), dgemv 'n' order of
    //    dmrgg.f90:986-991: y = 0; y = y + w(j) * a(., j, .) for ascending j.  One thread per entry (the sum is a
    //    sequential chain by definition); the operands are staged chunk by chunk by the whole CTA with
    //    independent loads in flight.
    const int ncol = grow_col ? r0 : 0, nrow = grow_row ? r1 : 0;
    int JC = (stage_doubles - 8) / (ncol + nrow + 1) - 1;
    if (JC > n) JC = n;
    const int half = blockDim.x >> 1;
    const bool col_thread = (int)threadIdx.x < ncol, row_thread = (int)threadIdx.x >= half && (int)threadIdx.x - half < nrow;
    double y = 0.0;
    for (int j0 = 0; j0 < n; j0 += JC) {
        const int jc = min(JC, n - j0);
        double* WS = stg; double* Sc = WS + jc; double* Sr = Sc + ncol * jc;      // WS[jj], Sc[i + ncol*jj], Sr[k*(jc+1) + jj]
        for (int jj = threadIdx.x; jj < jc; jj += blockDim.x) WS[jj] = use_weights ? w[j0 + jj] : 1.0;
        const double* ac = a + (i64)R * n * e1 + (i64)R * j0;
        constexpr int MB = 8;                                            // loads in flight per thread
        for (int i0 = 0; i0 < ncol; i0 += 32) {
            const int i = min(i0 + lane, ncol - 1);
            for (int jb = wid; jb < jc; jb += MB * nw) {
                double vv[MB];
#pragma unroll
                for (int u = 0; u < MB; ++u) vv[u] = ac[i + (i64)R * min(jb + u * nw, jc - 1)];
#pragma unroll
                for (int u = 0; u < MB; ++u) { const int jj = jb + u * nw; if (jj < jc && i0 + lane < ncol) Sc[i + ncol * jj] = vv[u]; }
            }
        }
        const double* ar = a + e0 + (i64)R * j0;
        for (int k = wid; k < nrow; k += nw) {
            const double* ark = ar + (i64)R * n * k;
            double* srk = Sr + k * (jc + 1);
            for (int jb = lane; jb < jc; jb += MB * 32) {
                double vv[MB];
#pragma unroll
                for (int u = 0; u < MB; ++u) vv[u] = ark[(i64)R * min(jb + u * 32, jc - 1)];
#pragma unroll
                for (int u = 0; u < MB; ++u) { const int jj = jb + u * 32; if (jj < jc) srk[jj] = vv[u]; }
            }
        }
        __syncthreads();
        tl_mark0(P, 61);
        if (col_thread) {
            const double* sc = Sc + threadIdx.x;
            if (use_weights) {
#pragma unroll 8
                for (int jj = 0; jj < jc; ++jj) y = y + WS[jj] * sc[ncol * jj];
            } else {
#pragma unroll 8
                for (int jj = 0; jj < jc; ++jj) y = y + sc[ncol * jj];
            }
        } else if (row_thread) {
            const double* sr = Sr + (threadIdx.x - half) * (jc + 1);
            if (use_weights) {
#pragma unroll 8
                for (int jj = 0; jj < jc; ++jj) y = y + WS[jj] * sr[jj];
            } else {
#pragma unroll 8
                for (int jj = 0; jj < jc; ++jj) y = y + sr[jj];
            }
        }
        __syncthreads();
        tl_mark0(P, 62);
    }
    if (col_thread) rawcol[threadIdx.x] = y;
    if (row_thread) rawrow[threadIdx.x - half] = y;
    __syncthreads();
    // 2. d2_luar: new column as a wavefront chain (warp 0); new row element of every old column (the other warps)
    if (grow_col && wid == 0) {
        double yy[MAXRPL];
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; yy[t] = (sidx < r0) ? rawcol[sidx] : 0.0; }
        warp_luar(yy, r0, GSm{TL, r0});
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; if (sidx < r0) Y[sidx + ld * e1] = yy[t]; }
    }
    if (grow_row && wid >= 1) {
        for (int k = (int)threadIdx.x - 32; k < e1; k += blockDim.x - 32) {
            double tmp = 0.0;
            const double* yk = YS + e0 * k;
#pragma unroll 8
            for (int u = 0; u < e0; ++u) tmp = tmp + yk[u] * TL[u * r0 + e0];
            Y[e0 + ld * k] = (e0 > 0) ? rawrow[k] + (-tmp) : rawrow[k];
        }
    }
    __syncthreads();
    tl_mark0(P, 63);
    // 3. d2_lual (not for the last core): new row as a wavefront chain; new column element of every old row
    if (p < P.d) {
        if (grow_row && wid == 0) {
            double yy[MAXRPL];
#pragma unroll
            for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; yy[t] = (c < r1) ? Y[e0 + ld * c] : 0.0; }
            warp_lual(yy, r1, GSm{TR, r1}, DSm{DI});
#pragma unroll
            for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; if (c < r1) Z[e0 + ld * c] = yy[t]; }
        }
        if (grow_col && wid >= 1) {
            for (int i = (int)threadIdx.x - 32; i < e0; i += blockDim.x - 32) {
                double val = Y[i + ld * e1];
#pragma unroll 8
                for (int u = 0; u < e1; ++u) val = val + (-TR[u * r1 + e1]) * ZS[i + e0 * u];
                val = DI[e1] * val;
                Z[i + ld * e1] = val;
            }
        }
    } else {
        if (grow_col) for (int i = threadIdx.x; i < r0; i += blockDim.x) Z[i + ld * e1] = Y[i + ld * e1];
        if (grow_row) for (int k = threadIdx.x; k < r1; k += blockDim.x) Z[e0 + ld * k] = Y[e0 + ld * k];
    }
    if (threadIdx.x == 0) { P.qext[2 * p] = r0; P.qext[2 * p + 1] = r1; }
    tl_mark0(P, 64);
}

// ----------------------------------------------------------------------------
// warp-per-fiber versions of the factor extensions (k_update_nbr, k_exchange_extend): the packed LU is staged
// transposed in shared memory once per CTA, every warp runs one mode index as a wavefront.
// blockIdx.z = 0: d2_luar part, 1: d2_lual part.
// ----------------------------------------------------------------------------
struct ExtJob {               // one batch of independent recurrences sharing one packed LU
    const double* g; int r;   // packed LU and its size
    int count;                // number of independent chains (mode indices)
    const double* src; i64 src_chain, src_elem;   // chain x, element s : src[x*src_chain + s*src_elem]
    double* dst; i64 dst_chain, dst_elem;
};
__device__ __forceinline__ void run_ext_luar(const ExtJob& J, double* sm) {
    stage_luar(J.g, J.r, sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int x = blockIdx.x * nw + wid; x < J.count; x += gridDim.x * nw) {
        double y[MAXRPL];
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; y[t] = (sidx < J.r) ? J.src[x * J.src_chain + sidx * J.src_elem] : 0.0; }
        warp_luar(y, J.r, GSm{sm, J.r});
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int sidx = lane + 32 * t; if (sidx < J.r) J.dst[x * J.dst_chain + sidx * J.dst_elem] = y[t]; }
    }
}
__device__ __forceinline__ void run_ext_lual(const ExtJob& J, double* sm) {
    double* di = sm + J.r * J.r;
    stage_lual(J.g, J.r, sm, di);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int x = blockIdx.x * nw + wid; x < J.count; x += gridDim.x * nw) {
        double y[MAXRPL];
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; y[t] = (c < J.r) ? J.src[x * J.src_chain + c * J.src_elem] : 0.0; }
        warp_lual(y, J.r, GSm{sm, J.r}, DSm{di});
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; if (c < J.r) J.dst[x * J.dst_chain + c * J.dst_elem] = y[t]; }
    }
}
// neighbour factors after an accepted pivot (dmrgg.f90:715-749)
__global__ void k_update_nbr_w(DevPlan P, int dir, int pp) {
    tl_stamp(P, 24);
    extern __shared__ double smem[];
    if (P.ctrl->ready) return;
    const int v = P.v0 + blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active || !P.st[v].upd) return;
    const int lo = P.own[v], hi = P.own[v + 1], t = D.r1;
    if (blockIdx.z == 0) {
        if (D.p <= lo || D.r0 < 1) return;
        ExtJob J;   // chains j (n1 of them), elements s < r0: src acol1(s,j), dst rowT(p)(s, j, q = t)
        J.g = P.inv + (i64)(D.p - 1) * P.Rmax * P.Rmax; J.r = D.r0; J.count = D.n1;
        J.src = P.acol1 + (i64)v * P.Rmax * P.nmax; J.src_chain = D.r0; J.src_elem = 1;
        J.dst = P.rowT + P.coreOff[D.p] + (i64)D.n1 * t; J.dst_chain = 1; J.dst_elem = (i64)D.n1 * P.Rmax;
        run_ext_luar(J, smem);
    } else {
        if (D.p >= hi - 1 || D.r2 < 1) return;
        ExtJob J;   // chains k (n2), elements c < r2: src arow1(k,c), dst col(p+1)(t, k, c)
        J.g = P.inv + (i64)(D.p + 1) * P.Rmax * P.Rmax; J.r = D.r2; J.count = D.n2;
        J.src = P.arow1 + (i64)v * P.Rmax * P.nmax; J.src_chain = 1; J.src_elem = D.n2;
        J.dst = P.col + P.coreOff[D.p + 1] + t; J.dst_chain = P.Rmax; J.dst_elem = (i64)P.Rmax * D.n2;
        run_ext_lual(J, smem);
    }
}
// factor extensions of the neighbour exchange (dmrgg.f90:939-951, dmrggmp.f90:616-626)
__global__ void k_exchange_extend_w(DevPlan P) {
    tl_stamp(P, 25);
    extern __shared__ double smem[];
    if (P.ctrl->ready) return;
    const int b = first_boundary(P) + blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    const int nc = P.n[c];
    const double* argc = P.arg + P.coreOff[c];
    if (blockIdx.z == 0) {
        if (!(rc > rcs) || !own_vrank(P, b)) return;
        ExtJob J;   // LEFT receiver: chains k, elements s < rc1: src arg(c)(s,k,rc), dst rowT(c)(s,k,rc)
        J.g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax; J.r = rc1; J.count = nc;
        J.src = argc + (i64)P.Rmax * nc * (rc - 1); J.src_chain = P.Rmax; J.src_elem = 1;
        J.dst = P.rowT + P.coreOff[c] + (i64)nc * (rc - 1); J.dst_chain = 1; J.dst_elem = (i64)nc * P.Rmax;
        run_ext_luar(J, smem);
    } else {
        if (!(rc1 > rc1s) || !own_vrank(P, b + 1)) return;
        ExtJob J;   // RIGHT receiver: chains j, elements cc < rc: src arg(c)(rc1,j,cc), dst col(c)(rc1,j,cc)
        J.g = P.inv + (i64)c * P.Rmax * P.Rmax; J.r = rc; J.count = nc;
        J.src = argc + (rc1 - 1); J.src_chain = P.Rmax; J.src_elem = (i64)P.Rmax * nc;
        J.dst = P.col + P.coreOff[c] + (rc1 - 1); J.dst_chain = P.Rmax; J.dst_elem = (i64)P.Rmax * nc;
        run_ext_lual(J, smem);
    }
}
// finalisation with wavefronts: d2_luar over the n*r1 columns, then d2_lual over the r0*n rows of every core
__global__ void k_lua_r_w(DevPlan P) {
    tl_stamp(P, 26);
    extern __shared__ double smem[];
    const int p = P.c_lo + blockIdx.y;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    if (r0 < 2) return;
    ExtJob J;
    J.g = P.inv + (i64)(p - 1) * P.Rmax * P.Rmax; J.r = r0; J.count = n * r1;
    J.src = P.arg + P.coreOff[p]; J.src_chain = P.Rmax; J.src_elem = 1;
    J.dst = P.arg + P.coreOff[p]; J.dst_chain = P.Rmax; J.dst_elem = 1;
    run_ext_luar(J, smem);
}
__global__ void k_lua_l_w(DevPlan P) {
    tl_stamp(P, 27);
    extern __shared__ double smem[];
    const int p = P.c_lo + blockIdx.y;
    if (p >= P.d) return;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    // chains are the rows (i,j): x = i + r0*j lives at i + Rmax*j -> two-level stride; run per j with chain index i
    double* di = smem + r1 * r1;
    stage_lual(P.inv + (i64)p * P.Rmax * P.Rmax, r1, smem, di);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* a = P.arg + P.coreOff[p];
    const i64 ys = (i64)P.Rmax * n;
    for (int x = blockIdx.x * nw + wid; x < r0 * n; x += gridDim.x * nw) {
        const int j = x / r0, i = x - j * r0;
        double* base = a + i + (i64)P.Rmax * j;
        double y[MAXRPL];
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; y[t] = (c < r1) ? base[c * ys] : 0.0; }
        warp_lual(y, r1, GSm{smem, r1}, DSm{di});
#pragma unroll
        for (int t = 0; t < MAXRPL; ++t) { int c = lane + 32 * t; if (c < r1) base[c * ys] = y[t]; }
    }
}

// Finalisation in ONE kernel (dtt_lua, dmrgg.f90:1169-1258): a warp owns the r0 x r1 slab arg(p)(:, j, :) of one mode index.  The slab
// is staged in shared memory with coalesced loads (the left index is contiguous in HBM), d2_luar (lr.f90:124-137) runs down its
// columns with one THREAD per column, d2_lual (lr.f90:139-154) along its rows with one thread per row; both packed LUs sit in
// shared memory and are read as broadcasts.  Every output accumulates the same terms in the same ascending order as the
// reference and as warp_luar / warp_lual -- four rows at a time share the pass over the finished prefix, which only interleaves
// independent chains -- so the cores come out bit-identical to k_lua_r_w + k_lua_l_w, at 1/32 of their dependent steps per
// element (a thread walks its chain out of shared memory; the wavefront kernels spend a whole warp and a shuffle per step).
// With `pack` the finished slab also lands in the caller-bound packed copy (k_pack_all's layout), saving its pass over the cores.
constexpr int LF_WARPS = 4;
__host__ __device__ __forceinline__ size_t lua_fused_smem(int Rmax) {
    return ((size_t)2 * Rmax * Rmax + Rmax + (size_t)LF_WARPS * (Rmax | 1) * Rmax) * sizeof(double);
}
__global__ void __launch_bounds__(32 * LF_WARPS) k_lua_fused(DevPlan P, double* pack) {
    tl_stamp(P, 37);
    extern __shared__ double smem[];
    const int p = P.c_lo + blockIdx.y;
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const bool do_r = r0 >= 2, do_l = p < P.d;
    if (!do_r && !do_l && !pack) return;
    double* TL = smem;                         // TL[u*r0 + s] = g_left(s,u), u < s      (stage_luar)
    double* TR = TL + P.Rmax * P.Rmax;         // TR[u*r1 + c] = g_right(c,u), u < c     (stage_lual)
    double* DI = TR + P.Rmax * P.Rmax;         // 1 / pivot(c)
    const int ld = r0 | 1;                     // odd leading dimension: a thread per column walks its column without bank conflicts
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* S = DI + P.Rmax + (size_t)wid * (P.Rmax | 1) * P.Rmax;
    if (do_r) stage_luar(P.inv + (i64)(p - 1) * P.Rmax * P.Rmax, r0, TL);
    if (do_l) stage_lual(P.inv + (i64)p * P.Rmax * P.Rmax, r1, TR, DI);
    __syncthreads();
    i64 poff = 0;
    if (pack) for (int k = P.c_lo; k < p; ++k) poff += (i64)P.rk[k - 1] * P.n[k] * P.rk[k];
    double* a = P.arg + P.coreOff[p];
    const i64 ys = (i64)P.Rmax * n;
    for (int j = blockIdx.x * LF_WARPS + wid; j < n; j += gridDim.x * LF_WARPS) {
        double* base = a + (i64)P.Rmax * j;
#pragma unroll 8
        for (int c = 0; c < r1; ++c)
            for (int i = lane; i < r0; i += 32) S[c * ld + i] = base[i + c * ys];
        __syncwarp();
        if (do_r) {
            for (int c = lane; c < r1; c += 32) {
                double* y = S + c * ld;
                for (int s0 = 1; s0 < r0; s0 += 4) {
                    const int sb = min(s0 + 1, r0 - 1), sc = min(s0 + 2, r0 - 1), sd = min(s0 + 3, r0 - 1);
                    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll 4
                    for (int u = 0; u < s0; ++u) {
                        const double yu = y[u];
                        const double* g = TL + u * r0;
                        t0 = t0 + yu * g[s0]; t1 = t1 + yu * g[sb]; t2 = t2 + yu * g[sc]; t3 = t3 + yu * g[sd];
                    }
                    const double y0 = y[s0] + (-t0);
                    y[s0] = y0;
                    if (s0 + 1 < r0) {
                        t1 = t1 + y0 * TL[s0 * r0 + s0 + 1];
                        const double y1 = y[s0 + 1] + (-t1);
                        y[s0 + 1] = y1;
                        if (s0 + 2 < r0) {
                            t2 = t2 + y0 * TL[s0 * r0 + s0 + 2];
                            t2 = t2 + y1 * TL[(s0 + 1) * r0 + s0 + 2];
                            const double y2 = y[s0 + 2] + (-t2);
                            y[s0 + 2] = y2;
                            if (s0 + 3 < r0) {
                                t3 = t3 + y0 * TL[s0 * r0 + s0 + 3];
                                t3 = t3 + y1 * TL[(s0 + 1) * r0 + s0 + 3];
                                t3 = t3 + y2 * TL[(s0 + 2) * r0 + s0 + 3];
                                y[s0 + 3] = y[s0 + 3] + (-t3);
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (do_l) {
            for (int i = lane; i < r0; i += 32) {
                double* y = S + i;          // y(c) = y[c*ld]
                y[0] = DI[0] * y[0];
                for (int c0 = 1; c0 < r1; c0 += 4) {
                    const int cb = min(c0 + 1, r1 - 1), cc = min(c0 + 2, r1 - 1), cd = min(c0 + 3, r1 - 1);
                    double v0 = y[c0 * ld], v1 = y[cb * ld], v2 = y[cc * ld], v3 = y[cd * ld];
#pragma unroll 4
                    for (int u = 0; u < c0; ++u) {
                        const double yu = y[u * ld];
                        const double* g = TR + u * r1;
                        v0 = v0 + (-g[c0]) * yu; v1 = v1 + (-g[cb]) * yu; v2 = v2 + (-g[cc]) * yu; v3 = v3 + (-g[cd]) * yu;
                    }
                    v0 = DI[c0] * v0;
                    y[c0 * ld] = v0;
                    if (c0 + 1 < r1) {
                        v1 = v1 + (-TR[c0 * r1 + c0 + 1]) * v0;
                        v1 = DI[c0 + 1] * v1;
                        y[(c0 + 1) * ld] = v1;
                        if (c0 + 2 < r1) {
                            v2 = v2 + (-TR[c0 * r1 + c0 + 2]) * v0;
                            v2 = v2 + (-TR[(c0 + 1) * r1 + c0 + 2]) * v1;
                            v2 = DI[c0 + 2] * v2;
                            y[(c0 + 2) * ld] = v2;
                            if (c0 + 3 < r1) {
                                v3 = v3 + (-TR[c0 * r1 + c0 + 3]) * v0;
                                v3 = v3 + (-TR[(c0 + 1) * r1 + c0 + 3]) * v1;
                                v3 = v3 + (-TR[(c0 + 2) * r1 + c0 + 3]) * v2;
                                v3 = DI[c0 + 3] * v3;
                                y[(c0 + 3) * ld] = v3;
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (do_r || do_l) {
#pragma unroll 8
            for (int c = 0; c < r1; ++c)
                for (int i = lane; i < r0; i += 32) base[i + c * ys] = S[c * ld + i];
        }
        if (pack) {
            double* o = pack + poff + (i64)r0 * j;
            const i64 os = (i64)r0 * n;
#pragma unroll 8
            for (int c = 0; c < r1; ++c)
                for (int i = lane; i < r0; i += 32) o[i + c * os] = S[c * ld + i];
        }
        __syncwarp();
    }
}

// factors of the initial cross (dmrgg.f90:234-248): inv(p)(1) = pivot, col(p) = arg(p)/pivot (d2_lual, r = 1),
// row(p) = arg(p) (d2_luar with r = 1 is the identity).  blockIdx.y = core - 1.
__global__ void k_init_factors(DevPlan P) {
    tl_stamp(P, 28);
    const int p = blockIdx.y + 1;
    const int n = P.n[p];
    const double* a = P.arg + P.coreOff[p];
    double* c = P.col + P.coreOff[p];
    double* r = P.rowT + P.coreOff[p];
    double sc = 1.0, pivot = 1.0;
    if (p < P.d) {
        int jp = P.vip[((i64)p * P.Rmax + 0) * 4 + 1];     // ind(p) of the initial cross
        pivot = a[(i64)P.Rmax * (jp - 1)];
        sc = 1.0 / pivot;
    }
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double x = a[(i64)P.Rmax * j];
        c[(i64)P.Rmax * j] = (p < P.d) ? sc * x : x;
        r[j] = x;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.inv[(i64)p * P.Rmax * P.Rmax] = pivot;
        if (p == 1) P.inv[0] = 1.0;
    }
}

// =============================================================================
// Core blocks over several processes (one per GPU).  The reference moves, after every sweep, the pivot tape
// (dmrgg.f90:763-850), three scalars (:852-870) and the new column / row of every shared core
// (:872-958, dmrggmp.f90:572-629) between MPI ranks, and dtt_lua hands inv(last bond) to the right (:1209-1246).
// Here every process packs (1) the visit records + scalar state of its virtual ranks into an all-gather mailbox and
// (2) the boundary slabs for its two neighbour processes; NCCL moves them (ttc_engine.cu); the kernels below replay
// the foreign pivots into the replicated index tables and drop the slabs into place.  After that the single-process
// exchange kernels (corner, factor extension) run unchanged on the boundaries this process touches.
// =============================================================================
constexpr int VO_WORDS = sizeof(VisitOut) / 8, VS_WORDS = sizeof(VState) / 8;
__host__ __device__ __forceinline__ int mb1_slot_words(int maxnb) { return maxnb * VO_WORDS + VS_WORDS; }
__host__ __device__ __forceinline__ int mb2_slot_doubles(int Rmax) { return Rmax * Rmax + 4; }

// grid: nv + 2 CTAs.  CTA b < nv: mailbox slot of virtual rank v0+b.  CTA nv: slab for the left process.  CTA nv+1: right.
// ----------------------------------------------------------------------------
// Peer-memory exchange: the pack kernels STORE straight into the other ranks' windows over NVLink (mapped with CUDA
// IPC), followed by a system-scope fence and one sequence-number store per destination; the unpack kernels spin on
// their own window's flags.  One process drives one GPU, so every spinning kernel has its producer running on another
// GPU; a bounded spin turns a lost peer into an error instead of a hang.
//   phase 1: mailbox slots of this rank's partitions -> every rank; column slab -> left neighbour's nb_recv_r area;
//            row | inv -> right neighbour's nb_recv_l area.          phase 2 / 3: chain-product slots -> every rank.
// ----------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mp_seq(const DevPlan& P, int phase) {
    if (phase == 3) return P.ctrl->quad_serial + 1ULL;
    return P.ctrl->run_serial * 65536ULL + (unsigned long long)P.ctrl->it;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// epilogue of a packing kernel (every thread of every CTA calls it): once the whole grid's stores are fenced, the last CTA
// stores this phase's sequence number into every rank's flag slot for this source
__device__ __forceinline__ void mp_publish(const DevPlan& P, int phase) {
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(P.tickets + P.P, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) P.tickets[P.P] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    const unsigned long long seq = mp_seq(P, phase);
    for (int q = threadIdx.x; q < P.nproc; q += blockDim.x)
        st_release_sys((unsigned long long*)(P.peer_win[q] + P.win_flg) + (long long)(phase - 1) * P.nproc + P.prank, seq);
    if (phase == 3 && threadIdx.x == 0) P.ctrl->quad_serial = seq;
}
__global__ void k_mp_pack1(DevPlan P) {
    tl_stamp(P, 29);
    if (P.ctrl->ready) return;
    const int it = P.ctrl->it;
    const int b = blockIdx.x;
    const bool p2p = P.peer_win != nullptr;
    const int sw = mb1_slot_words(P.maxnb);
    if (b < P.nv) {
        const int v = P.v0 + b;
        for (int g = 0; g < (p2p ? P.nproc : 1); ++g) {
            unsigned long long* slot = p2p ? (unsigned long long*)(P.peer_win[g] + P.win_mb1) + ((i64)P.prank * P.vper + b) * sw
                                           : P.mb1_send + (i64)b * sw;
            for (int x = threadIdx.x; x < P.maxnb * VO_WORDS; x += blockDim.x) {
                const int pp = x / VO_WORDS, w = x - pp * VO_WORDS;
                slot[x] = ((const unsigned long long*)&P.vlog[((i64)(it - 1) * P.maxnb + pp) * P.P + v])[w];
            }
            for (int x = threadIdx.x; x < VS_WORDS; x += blockDim.x)
                slot[P.maxnb * VO_WORDS + x] = ((const unsigned long long*)&P.st[v])[x];
        }
    } else if (b == P.nv) {
        if (P.v0 > 0) {
            const int c = P.own[P.v0];                       // shared with the left process; our bond c
            if (P.rk[c] > P.rks[c]) {
                const int n = P.n[c];
                const double* slab = P.arg + P.coreOff[c] + (i64)P.Rmax * n * P.rks[c];   // new slice, contiguous
                double* dst = p2p ? (double*)(P.peer_win[P.prank - 1] + P.win_nbr) : P.nb_send_l;   // the left rank receives "from the right"
                for (int x = threadIdx.x; x < P.Rmax * n; x += blockDim.x) dst[x] = slab[x];
            }
        }
    } else {
        if (P.v0 + P.nv < P.P) {
            const int c = P.own[P.v0 + P.nv];                // shared with the right process; our bond c-1
            const int n = P.n[c];
            const double* g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax;
            double* dst = p2p ? (double*)(P.peer_win[P.prank + 1] + P.win_nbl) : P.nb_send_r;       // the right rank receives "from the left"
            double* ginv = dst + (i64)P.nmax * P.Rmax;
            for (int x = threadIdx.x; x < P.Rmax * P.Rmax; x += blockDim.x) ginv[x] = g[x];
            if (P.rk[c - 1] > P.rks[c - 1]) {
                const double* a = P.arg + P.coreOff[c] + P.rks[c - 1];                    // new row t, stride Rmax
                for (int x = threadIdx.x; x < n * P.Rmax; x += blockDim.x) dst[x] = a[(i64)P.Rmax * x];
            }
        }
    }
    if (p2p) mp_publish(P, 1);
}
// every consumer CTA calls this first: wait until all ranks' pushes of this phase have landed in the local window
__device__ __forceinline__ void mp_wait(const DevPlan& P, int phase) {
    if (!P.peer_win) return;                        // NCCL transport: the stream orders the receive
    if (threadIdx.x == 0) {
        // phase 3 is consumed after this rank's own push already advanced quad_serial
        const unsigned long long seq = (phase == 3) ? P.ctrl->quad_serial : mp_seq(P, phase);
        unsigned long long t0 = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (int q = 0; q < P.nproc; ++q) {
            const unsigned long long* f = P.win_flags + (long long)(phase - 1) * P.nproc + q;
            while (ld_acquire_sys(f) < seq) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ULL) { P.ctrl->error = 3; break; }     // 4 s: a peer is gone
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}
// grid: P CTAs, CTA v handles foreign virtual rank v: state + visit records, then the replay of its accepted pivots
// (the index-set half of k_accept) in visit order.
__global__ void k_mp_unpack1(DevPlan P) {
    tl_stamp(P, 30);
    if (P.ctrl->ready) return;
    mp_wait(P, 1);
    const int v = blockIdx.x;
    if (own_vrank(P, v)) return;
    const int it = P.ctrl->it;
    int g = 0;
    while (g + 1 < P.nproc && proc_v0(P.P, P.nproc, g + 1) <= v) ++g;
    const unsigned long long* slot = P.mb1_recv + ((i64)g * P.vper + (v - proc_v0(P.P, P.nproc, g))) * mb1_slot_words(P.maxnb);
    for (int x = threadIdx.x; x < P.maxnb * VO_WORDS; x += blockDim.x) {
        const int pp = x / VO_WORDS, w = x - pp * VO_WORDS;
        ((unsigned long long*)&P.vlog[((i64)(it - 1) * P.maxnb + pp) * P.P + v])[w] = slot[x];
    }
    for (int x = threadIdx.x; x < VS_WORDS; x += blockDim.x) ((unsigned long long*)&P.st[v])[x] = slot[P.maxnb * VO_WORDS + x];
    __syncthreads();
    for (int pp = 0; pp < P.maxnb; ++pp) {
        const VisitOut& O = P.vlog[((i64)(it - 1) * P.maxnb + pp) * P.P + v];
        if (!(O.active && O.upd)) continue;               // uniform over the CTA
        const int p = O.bond, t = P.rk[p];
        if (t >= P.Rmax) { if (threadIdx.x == 0) P.ctrl->error = 1; continue; }
        const int ii = O.ii, jj = O.jj, kk = O.kk, qq = O.qq;
        if (threadIdx.x == 0) { int* vp = P.vip + ((i64)p * P.Rmax + t) * 4; vp[0] = ii; vp[1] = jj; vp[2] = kk; vp[3] = qq; }
        int* Lp = P.Lidx + P.offL[p];
        const int* Lm = P.Lidx + P.offL[p - 1];
        for (int pos = threadIdx.x; pos < p; pos += blockDim.x)
            Lp[(i64)pos * P.Rmax + t] = (pos < p - 1) ? Lm[(i64)pos * P.Rmax + (ii - 1)] : jj;
        int* Rp = P.Ridx + P.offR[p];
        const int* Rn = P.Ridx + P.offR[p + 1];
        for (int pos = threadIdx.x; pos < P.d - p; pos += blockDim.x)
            Rp[(i64)pos * P.Rmax + t] = (pos == 0) ? kk : Rn[(i64)(pos - 1) * P.Rmax + (qq - 1)];
        __syncthreads();
        if (threadIdx.x == 0) P.rk[p] = t + 1;
        __syncthreads();
    }
}
// grid: (NB, 2).  y = 0: what the left process sent (new row of core own[v0] + inv of its last bond);
// y = 1: what the right process sent (new column slab of core own[v0+nv]).  Runs after k_mp_unpack1 (ranks replayed).
__global__ void k_mp_unpack1b(DevPlan P) {
    tl_stamp(P, 31);
    if (P.ctrl->ready) return;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && P.P > 1) {
        // MPI_ALLREDUCE(MAX) of (amax, pivotmax, -pivotmin) (dmrgg.f90:852-870) over the gathered states (k_allreduce)
        double c1 = P.st[0].amax, c2 = P.st[0].pivotmax, c3 = (P.st[0].pivotmin > 0.0) ? -P.st[0].pivotmin : -999e9;
        for (int v = 1; v < P.P; ++v) {
            c1 = fmax(c1, P.st[v].amax); c2 = fmax(c2, P.st[v].pivotmax);
            c3 = fmax(c3, (P.st[v].pivotmin > 0.0) ? -P.st[v].pivotmin : -999e9);
        }
        for (int v = 0; v < P.P; ++v) {
            P.st[v].amax = c1; P.st[v].pivotmax = c2; P.st[v].pivotmin = (-c3 == 999e9) ? -1.0 : -c3;
        }
    }
    if (blockIdx.y == 0) {
        if (P.v0 == 0) return;
        const int c = P.own[P.v0];
        double* g = P.inv + (i64)(c - 1) * P.Rmax * P.Rmax;
        const double* ginv = P.nb_recv_l + (i64)P.nmax * P.Rmax;
        for (int x = tid; x < P.Rmax * P.Rmax; x += nth) g[x] = ginv[x];
        if (!(P.rk[c - 1] > P.rks[c - 1])) return;
        const int n = P.n[c], t = P.rks[c - 1], rq = P.rks[c];       // the sender saw bond c at its sweep-start rank
        double* a = P.arg + P.coreOff[c] + t;
        for (int x = tid; x < n * rq; x += nth) a[(i64)P.Rmax * x] = P.nb_recv_l[x];
    } else {
        if (P.v0 + P.nv >= P.P) return;
        const int c = P.own[P.v0 + P.nv];
        if (!(P.rk[c] > P.rks[c])) return;
        const int n = P.n[c], ri = P.rks[c - 1];                      // the sender saw bond c-1 at its sweep-start rank
        double* slab = P.arg + P.coreOff[c] + (i64)P.Rmax * n * P.rks[c];
        for (int x = tid; x < P.Rmax * n; x += nth) { if (x % P.Rmax < ri) slab[x] = P.nb_recv_r[x]; }
    }
}
// phase 2 (after the corner evaluations and the per-rank quadrature chains): chain products, amax, neval, error flag
// final != 0: the collective ttc_quad after the run (not gated by the ready flag; only the chain products travel)
__global__ void k_mp_pack2(DevPlan P, int final) {
    tl_stamp(P, 32);
    if (!final && P.ctrl->ready) return;
    const int b = blockIdx.x, v = P.v0 + b;
    const int msz = P.Rmax * P.Rmax;
    const bool p2p = P.peer_win != nullptr;
    const double* ch = P.chain + (i64)v * msz;
    for (int g = 0; g < (p2p ? P.nproc : 1); ++g) {
        double* slot = p2p ? (double*)(P.peer_win[g] + P.win_mb2) + ((i64)P.prank * P.vper + b) * mb2_slot_doubles(P.Rmax)
                           : P.mb2_send + (i64)b * mb2_slot_doubles(P.Rmax);
        for (int x = threadIdx.x; x < msz; x += blockDim.x) slot[x] = ch[x];
        if (threadIdx.x == 0) {
            slot[msz] = P.st[v].amax;
            slot[msz + 1] = __longlong_as_double(P.st[v].neval);
            slot[msz + 2] = __longlong_as_double((long long)P.ctrl->error);
            slot[msz + 3] = 0.0;
        }
    }
    if (p2p) mp_publish(P, final ? 3 : 2);
}
__global__ void k_mp_unpack2(DevPlan P, int final) {
    tl_stamp(P, 33);
    if (!final && P.ctrl->ready) return;
    mp_wait(P, final ? 3 : 2);
    const int v = blockIdx.x;
    if (own_vrank(P, v)) return;
    int g = 0;
    while (g + 1 < P.nproc && proc_v0(P.P, P.nproc, g + 1) <= v) ++g;
    const int msz = P.Rmax * P.Rmax;
    const double* slot = P.mb2_recv + ((i64)g * P.vper + (v - proc_v0(P.P, P.nproc, g))) * mb2_slot_doubles(P.Rmax);
    double* ch = P.chain + (i64)v * msz;
    for (int x = threadIdx.x; x < msz; x += blockDim.x) ch[x] = slot[x];
    if (threadIdx.x == 0 && !final) {
        P.st[v].amax = slot[msz];
        P.st[v].neval = __double_as_longlong(slot[msz + 1]);
        if (__double_as_longlong(slot[msz + 2]) != 0) P.ctrl->error = 1;
    }
}

static const char* const tl_names[] = {"k_lot", "k_fiber", "k_superblock", "k_accept", "k_update_main", "k_update_nbr", "k_allreduce", "k_run_begin", "k_sweep_log", "k_exchange_corner", "k_exchange_extend", "k_quad_contract", "k_quad_lua", "k_quad_chain", "k_quad_tree", "k_lua_r", "k_lua_l", "k_pack_core", "k_init_search", "k_init_cross", "k_quad_contract_sm", "k_quad_lua_sm", "k_quad_chain_sm", "k_quad_tree_sm", "k_update_nbr_w", "k_exchange_extend_w", "k_lua_r_w", "k_lua_l_w", "k_init_factors", "k_mp_pack1", "k_mp_unpack1", "k_mp_unpack1b", "k_mp_pack2", "k_mp_unpack2", "k_quad_inc", "k_superblock_t", "k_sweeps", "k_lua_fused"};
}  // namespace ttc
