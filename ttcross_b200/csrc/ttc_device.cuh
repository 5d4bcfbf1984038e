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

namespace ttc {

typedef long long i64;

enum { KIND_ISING = 1, KIND_STDNORM = 4, KIND_MVN = 5 };
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
};
// exact restatement of lottery2's cumulative weights (rnd.f90:115-125) for 0/1 weights, see build_segments()
struct LotSeg { long long M; long long k; int c0; int J; int E; int pad; };
constexpr int MAXSEG = 128;

struct DevPlan {
    int d, P, Rmax, nmax, piv, kind, ising_id, nlotmax;
    int auxsm;             // doubles of dynamic shared memory reserved for the MVN matrix (0: read it from global)
    const int* n;          // n[1..d]; n[0] = n[d+1] = 1
    const int* own;        // own[0..P]
    const double* par;     // nodes | weights | ...
    const double* aux;     // MVN: mu | inv_cov | denom
    int* Lidx; int* Ridx; const i64* offL; const i64* offR;
    int* vip;              // [(d+1)][Rmax][4]
    int* rk; int* rks;     // ranks now / at sweep start, index 0..d
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
    unsigned long long seed;
    int dev_lottery;       // 1: lottery on the device (built-in uniform stream), 0: host fills `lot`
    int maxnb, maxsweeps;
    int has_accuracy; double accuracy;
    Ctrl* ctrl;
    VisitOut* vlog;        // [maxsweeps][maxnb][P]
    SweepOut* slog;        // [maxsweeps + 1]
    int* rklog;            // [maxsweeps + 1][d + 1]
};

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
// integrands
// ----------------------------------------------------------------------------
// test_crs_ising.f90:176-218.  Pure + - * / : bit-reproducible.
template <class Src>
__device__ double eval_ising(const DevPlan& P, const Src& s) {
    const int m = P.d;
    const int id = P.ising_id;
    const double* nodes = P.par - 1;
    const double* weights = P.par + P.n[1] - 1;
    double a = 0.0, b = 0.0, f;
    if (id == 2 || id == 3) {
        a = 1.0;
        if (m <= MAXD_LOCAL) {
            double x[MAXD_LOCAL];
            for (int j = 1; j <= m; ++j) x[j - 1] = nodes[s(j)];
            for (int i = 0; i <= m; ++i) {
                double uij = 1.0;
                for (int j = i + 1; j <= m; ++j) {
                    uij = uij * x[j - 1];
                    double t = (uij - 1.0) / (uij + 1.0);
                    a = a * (t * t);
                }
            }
        } else {
            for (int i = 0; i <= m; ++i) {
                double uij = 1.0;
                for (int j = i + 1; j <= m; ++j) {
                    uij = uij * nodes[s(j)];
                    double t = (uij - 1.0) / (uij + 1.0);
                    a = a * (t * t);
                }
            }
        }
    }
    if (id == 1 || id == 2) {
        double v = 1.0, w = 1.0, vk = 1.0, wk = 1.0;
        for (int i = 1; i <= m; ++i) {
            vk = vk * nodes[s(m - i + 1)];
            wk = wk * nodes[s(i)];
            v = v + vk;
            w = w + wk;
        }
        b = 1.0 / (v * w);
    }
    if (id == 1) f = 2 * b;
    else if (id == 2) f = 2 * a * b;
    else f = 2 * a;
    for (int i = 1; i <= m; ++i) f = f * weights[s(i)];
    return f;
}
// test_crs_stdnorm.f90:154-170
template <class Src>
__device__ double eval_stdnorm(const DevPlan& P, const Src& s) {
    double sum = 0.0;
    for (int i = 1; i <= P.d; ++i) { double x = P.par[s(i) - 1]; sum = sum + x * x; }
    return exp(-sum);
}
// lib/mvn_pdf.f90:63-83 (through test_crs_mvn.f90:156-172); A = inv_cov column-major, staged by the caller
template <class Src>
__device__ double eval_mvn(const DevPlan& P, const Src& s, const double* __restrict__ A /*d*d*/) {
    const int m = P.d;
    const double* mu = P.aux;
    const double denom = P.aux[m + (i64)m * m];
    double e = 0.0;
    if (m <= MAXD_LOCAL) {
        double diff[MAXD_LOCAL];
        for (int i = 0; i < m; ++i) diff[i] = P.par[s(i + 1) - 1] - mu[i];
        for (int i = 0; i < m; ++i) {
            const double di = diff[i];
            for (int j = 0; j < m; ++j) e = e + di * A[i + (i64)j * m] * diff[j];
        }
    } else {
        for (int i = 0; i < m; ++i) {
            const double di = P.par[s(i + 1) - 1] - mu[i];
            for (int j = 0; j < m; ++j) e = e + di * A[i + (i64)j * m] * (P.par[s(j + 1) - 1] - mu[j]);
        }
    }
    return exp(-0.5 * e) / denom;
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
template <int KIND, class Src>
__device__ __forceinline__ double eval_point(const DevPlan& P, const Src& s, const double* A) {
    if (KIND == KIND_ISING) return eval_ising(P, s);
    if (KIND == KIND_STDNORM) return eval_stdnorm(P, s);
    return eval_mvn(P, s, A);
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
__host__ __device__ inline int build_segments(int scol, LotSeg* seg) {
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
__host__ __device__ __forceinline__ double lot_T(const LotSeg* seg, int ns, int c) {   // T[c], 1 <= c <= scol
    int lo = 0, hi = ns - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (seg[mid].c0 <= c) lo = mid; else hi = mid - 1; }
    const LotSeg& g = seg[lo];
    return scalbn((double)(g.M + (long long)(c - g.c0) * g.k), g.E - 52);
}
// one draw: 1-based cell index in 1..m (zeros = sorted distinct 1-based zero-weight cells)
__host__ __device__ __forceinline__ int lot_draw(const LotSeg* seg, int ns, int scol, int m, const int* zeros, int nz, double y) {
    if (!(y < lot_T(seg, ns, scol))) return m;          // x(n) <= y  ->  n = m+1, clamped to m (rnd.f90:122)
    int lo = 1, hi = scol;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (y < lot_T(seg, ns, mid)) hi = mid; else lo = mid + 1; }
    int sidx = lo;                                       // the lo-th cell of non-zero weight
    for (int z = 0; z < nz; ++z) { if (zeros[z] <= sidx) ++sidx; else break; }
    return sidx;
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

// ----------------------------------------------------------------------------
// K1: lottery candidates (dmrgg.f90:447-484): evaluate, residual by sequential ddot, two argmaxes
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_lot(DevPlan P, int dir, int pp) {
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const double* A = stage_aux<KIND>(P, smem);
    const int nlot = D.r0 + D.n1 + D.n2 + D.r2;
    int* lot = P.lot + (i64)v * 4 * P.nlotmax;
    if (P.dev_lottery) {
        // every CTA rebuilds the (tiny) cumulative-weight description; each thread then draws its own candidates
        __shared__ LotSeg seg[2][MAXSEG];
        __shared__ int s_ns[2], s_nz[2];
        int* ibuf = (int*)(smem + P.auxsm);           // tmp[Rmax] | zeros_col[Rmax] | zeros_row[Rmax]
        int* tmp = ibuf; int* zc = ibuf + P.Rmax; int* zr = ibuf + 2 * P.Rmax;
        const int* vip_p = P.vip + (i64)D.p * P.Rmax * 4;
        const int m = D.r0 * D.n1, n = D.n2 * D.r2;
        lot_zeros(vip_p, D.r1, 0, D.r0, tmp, zc, &s_nz[0]);
        lot_zeros(vip_p, D.r1, 1, D.n2, tmp, zr, &s_nz[1]);
        if (threadIdx.x == 0)  s_ns[0] = build_segments(m - s_nz[0], seg[0]);
        if (threadIdx.x == 32 % blockDim.x && blockDim.x > 32) s_ns[1] = build_segments(n - s_nz[1], seg[1]);
        if (blockDim.x <= 32 && threadIdx.x == 0) s_ns[1] = build_segments(n - s_nz[1], seg[1]);
        __syncthreads();
        const unsigned long long k0 = P.st[v].rng_k;
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nlot; x += gridDim.x * blockDim.x) {
            double uc = stream_uniform(P.seed, v, k0 + (unsigned long long)x);
            double ur = stream_uniform(P.seed, v, k0 + (unsigned long long)(nlot + x));
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
        PointSrc s = bond_point(P, D.p, i, j, k, q);
        double f = eval_point<KIND>(P, s, A);
        const double* c = colp + (i - 1) + (i64)P.Rmax * (j - 1);
        const double* r = rowp + (k - 1) + (i64)D.n2 * (q - 1);
        double t = 0.0;
        for (int sidx = 0; sidx < D.r1; ++sidx) t = t + c[sidx * cs] * r[sidx * rs];
        double res = f - t;
        P.lraw[(i64)v * P.nlotmax + x] = f;
        P.lres[(i64)v * P.nlotmax + x] = res;
        amax_take(braw, f, x);
        amax_take(bres, res, x);
    }
    braw = amax_block(braw, shp);
    bres = amax_block(bres, shp);
    if (threadIdx.x == 0) {
        P.part[((i64)v * 2 + 0) * GMAX + blockIdx.x] = braw;
        P.part[((i64)v * 2 + 1) * GMAX + blockIdx.x] = bres;
    }
}

__device__ __forceinline__ Partial reduce_parts(const Partial* parts, int G, Partial* shp) {
    Partial a = amax_init();
    for (int x = threadIdx.x; x < G; x += blockDim.x) amax_merge(a, parts[x]);
    return amax_block(a, shp);
}

__global__ void k_lot_reduce(DevPlan P, int dir, int pp, int G) {
    if (P.ctrl->ready) return;
    __shared__ Partial shp[32];
    const int v = blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    Partial braw = reduce_parts(P.part + ((i64)v * 2 + 0) * GMAX, G, shp);
    Partial bres = reduce_parts(P.part + ((i64)v * 2 + 1) * GMAX, G, shp);
    if (threadIdx.x == 0) {
        VState& S = P.st[v];
        const int nlot = D.r0 + D.n1 + D.n2 + D.r2;
        const int* lot = P.lot + (i64)v * 4 * P.nlotmax;
        S.amax = fmax(S.amax, braw.absv);
        int x = (int)bres.idx;
        S.ii = lot[x]; S.jj = lot[P.nlotmax + x]; S.kk = lot[2 * P.nlotmax + x]; S.qq = lot[3 * P.nlotmax + x];
        S.pivot = bres.val;
        S.done = 0; S.havecol = 0; S.haverow = 0; S.crs = 0; S.upd = 0;
        S.neval += nlot;
        S.rng_k += 2ULL * (unsigned long long)nlot;     // one random_number(d(npnt,2)) call (rnd.f90:120)
    }
}

// ----------------------------------------------------------------------------
// K2: cross fibers of the rook search (dmrgg.f90:519-581) fused with their residuals
//   column fiber: acol1(i,j) = f(i,j,kk,qq); bcol1 = acol1 - col(p)(:,:,1:r) * row(p+1)(1:r,kk,qq)   [dgemv 'n' order]
//   row fiber   : arow1(k,q) = f(ii,jj,k,q); brow1 = arow1 - row(p+1)(1:r,:,:)^T col(p)(ii,jj,1:r)   [dgemv 't' order]
// mode 0: rook step (skipped when the visit is already `done`); mode 1: unconditional (piv = 0 and piv = -1 branches)
// ----------------------------------------------------------------------------
template <int KIND, int ISROW>
__global__ void k_fiber(DevPlan P, int dir, int pp, int mode) {
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    const VState& S = P.st[v];
    if (mode == 0 && S.done) return;
    // coefficient vector of the residual: row(p+1)(1:r1,kk,qq) or col(p)(ii,jj,1:r1)
    double* xs = smem;                         // [Rmax]
    double* Asm = smem + P.Rmax;
    const double* colp = P.col + P.coreOff[D.p];
    const double* rowp = P.rowT + P.coreOff[D.p + 1];
    const i64 cs = (i64)P.Rmax * D.n1;
    const i64 rs = (i64)D.n2 * P.Rmax;
    const int ii = S.ii, jj = S.jj, kk = S.kk, qq = S.qq;
    for (int s = threadIdx.x; s < D.r1; s += blockDim.x)
        xs[s] = ISROW ? colp[(ii - 1) + (i64)P.Rmax * (jj - 1) + s * cs] : rowp[(kk - 1) + (i64)D.n2 * (qq - 1) + s * rs];
    __syncthreads();
    const double* A = stage_aux<KIND>(P, Asm);
    const int count = ISROW ? D.n2 * D.r2 : D.r0 * D.n1;
    double* fa = (ISROW ? P.arow1 : P.acol1) + (i64)v * P.Rmax * P.nmax;
    double* fb = (ISROW ? P.brow1 : P.bcol1) + (i64)v * P.Rmax * P.nmax;
    Partial braw = amax_init(), bres = amax_init();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        double f, res;
        if (!ISROW) {
            int j = e / D.r0 + 1, i = e % D.r0 + 1;
            PointSrc s = bond_point(P, D.p, i, j, kk, qq);
            f = eval_point<KIND>(P, s, A);
            const double* c = colp + (i - 1) + (i64)P.Rmax * (j - 1);
            res = f;
            for (int sidx = 0; sidx < D.r1; ++sidx) res = res + (-xs[sidx]) * c[sidx * cs];
        } else {
            int q = e / D.n2 + 1, k = e % D.n2 + 1;
            PointSrc s = bond_point(P, D.p, ii, jj, k, q);
            f = eval_point<KIND>(P, s, A);
            const double* r = rowp + (k - 1) + (i64)D.n2 * (q - 1);
            double t = 0.0;
            for (int sidx = 0; sidx < D.r1; ++sidx) t = t + r[sidx * rs] * xs[sidx];
            res = f + (-t);
        }
        fa[e] = f;
        fb[e] = res;
        amax_take(braw, f, e);
        amax_take(bres, res, e);
    }
    braw = amax_block(braw, shp);
    bres = amax_block(bres, shp);
    if (threadIdx.x == 0) {
        P.part[((i64)v * 2 + 0) * GMAX + blockIdx.x] = braw;
        P.part[((i64)v * 2 + 1) * GMAX + blockIdx.x] = bres;
    }
}

// the scalar bookkeeping after a fiber (dmrgg.f90:527-547, 560-580)
template <int ISROW>
__global__ void k_fiber_reduce(DevPlan P, int dir, int pp, int mode, int G) {
    if (P.ctrl->ready) return;
    __shared__ Partial shp[32];
    const int v = blockIdx.y;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    VState& S = P.st[v];
    if (mode == 0 && S.done) return;
    Partial braw = reduce_parts(P.part + ((i64)v * 2 + 0) * GMAX, G, shp);
    Partial bres = reduce_parts(P.part + ((i64)v * 2 + 1) * GMAX, G, shp);
    if (threadIdx.x == 0) {
        const int count = ISROW ? D.n2 * D.r2 : D.r0 * D.n1;
        if (mode == 2) return;                       // piv = -1: fibers are slices of the superblock, nothing to account
        S.neval += count;
        if (mode == 1) { S.havecol = 1; S.haverow = 1; S.done = 1; return; }   // piv = 0 (dmrgg.f90:492-513)
        S.amax = fmax(S.amax, braw.absv);
        if (ISROW) S.haverow = 1; else S.havecol = 1;
        S.crs += 1;
        int done = S.havecol && S.haverow && (S.crs >= 2 * P.piv);
        if (!done) {
            int e = (int)bres.idx;
            if (!ISROW) {
                int j = e / D.r0 + 1, i = e % D.r0 + 1;
                done = S.havecol && S.haverow && (i == S.ii && j == S.jj);
                S.ii = i; S.jj = j;
            } else {
                int q = e / D.n2 + 1, k = e % D.n2 + 1;
                done = S.havecol && S.haverow && (k == S.kk && q == S.qq);
                S.kk = k; S.qq = q;
            }
            S.pivot = bres.val;
        }
        S.done = done;
    }
}

// ----------------------------------------------------------------------------
// K3: full-pivoting superblock (dmrgg.f90:341-396), fused: evaluate a(i,j,k,q), residual against col*row in
// dgemm order (K = r(p)), first-index argmax of |a| and of |b|.  STORE also writes `a` to HBM (HBM-bound variant).
// Tile: blockDim.x threads walk the linear index (i fastest) so stores and col reads are coalesced.
// ----------------------------------------------------------------------------
template <int KIND, int STORE>
__global__ void k_superblock(DevPlan P, int dir, int pp, int fixed_bond, int fixed_v, double* a_out) {
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int v = (fixed_bond > 0) ? fixed_v : blockIdx.y;
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
        PointSrc s = bond_point(P, D.p, i, j, k, q);
        double f = eval_point<KIND>(P, s, A);
        const double* c = colp + (i - 1) + (i64)P.Rmax * (j - 1);
        const double* r = rowp + (k - 1) + (i64)D.n2 * (q - 1);
        double res = f;
        for (int sidx = 0; sidx < D.r1; ++sidx) res = res + (-r[sidx * rs]) * c[sidx * cs];
        if (STORE) a_out[x] = f;
        amax_take(braw, f, x);
        amax_take(bres, res, x);
    }
    braw = amax_block(braw, shp);
    bres = amax_block(bres, shp);
    if (threadIdx.x == 0) {
        P.part[((i64)v * 2 + 0) * GMAX + blockIdx.x] = braw;
        P.part[((i64)v * 2 + 1) * GMAX + blockIdx.x] = bres;
    }
}

__global__ void k_superblock_reduce(DevPlan P, int dir, int pp, int G, int fixed_bond, int fixed_v, Partial* probe_out) {
    __shared__ Partial shp[32];
    const int v = (fixed_bond > 0) ? fixed_v : blockIdx.y;
    Dims D;
    if (fixed_bond > 0) {
        D.active = 1; D.p = fixed_bond; D.r0 = P.rk[D.p - 1]; D.r1 = P.rk[D.p]; D.r2 = P.rk[D.p + 1];
        D.n1 = P.n[D.p]; D.n2 = P.n[D.p + 1];
    } else {
        if (P.ctrl->ready) return;
        D = load_dims(P, v, dir, pp);
    }
    if (!D.active) return;
    Partial braw = reduce_parts(P.part + ((i64)v * 2 + 0) * GMAX, G, shp);
    Partial bres = reduce_parts(P.part + ((i64)v * 2 + 1) * GMAX, G, shp);
    if (threadIdx.x == 0) {
        if (probe_out) { probe_out[0] = braw; probe_out[1] = bres; return; }
        VState& S = P.st[v];
        const i64 m1 = (i64)D.r0 * D.n1;
        S.amax = fmax(S.amax, braw.absv);
        i64 x = bres.idx;
        i64 kq = x / m1; int ij = (int)(x - kq * m1);
        S.qq = (int)(kq / D.n2) + 1; S.kk = (int)(kq % D.n2) + 1;
        S.jj = ij / D.r0 + 1; S.ii = ij % D.r0 + 1;
        S.pivot = bres.val;
        S.done = 1; S.havecol = 1; S.haverow = 1; S.crs = 0; S.upd = 0;
        S.neval += m1 * D.n2 * D.r2;
    }
}

// ----------------------------------------------------------------------------
// K4: accept test and index-set update (dmrgg.f90:598-660)
// ----------------------------------------------------------------------------
__global__ void k_accept(DevPlan P, int it, int dir, int pp, double small_element, double small_pivot) {
    if (P.ctrl->ready) return;
    const int v = blockIdx.y;
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
    if (P.ctrl->ready) return;
    const int v = blockIdx.y;
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
}

// neighbour factors (dmrgg.f90:715-749): new column of row(p) through d2_luar(inv(p-1)), new row of col(p+1)
// through d2_lual(inv(p+1)).  One thread per mode index; the triangular recurrences are sequential by definition.
__global__ void k_update_nbr(DevPlan P, int dir, int pp) {
    if (P.ctrl->ready) return;
    const int v = blockIdx.y;
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

__global__ void k_end_visit(DevPlan P, int dir, int pp) {
    if (P.ctrl->ready) return;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= P.P) return;
    const Dims D = load_dims(P, v, dir, pp);
    if (!D.active) return;
    if (P.st[v].upd) P.rk[D.p] = D.r1 + 1;
}

// ----------------------------------------------------------------------------
// sweep begin / end
// ----------------------------------------------------------------------------
__global__ void k_sweep_begin(DevPlan P) {
    if (P.ctrl->ready) return;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x <= P.d) P.rks[x] = P.rk[x];
    if (x < P.P) { P.st[x].pivotmax = -1.0; P.st[x].pivotmin = -1.0; }
}
// MPI_ALLREDUCE(MAX) of (amax, pivotmax, -pivotmin) (dmrgg.f90:852-870); single thread, P is small
__global__ void k_allreduce(DevPlan P) {
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
__global__ void k_sweep_end(DevPlan P) {
    if (P.ctrl->ready) return;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    i64 ne = 0;
    for (int v = 0; v < P.P; ++v) ne += P.st[v].neval;
    P.sweep_out->neval = ne;
    P.sweep_out->amax = P.st[0].amax;
    P.sweep_out->pivotmax = P.st[0].pivotmax;
    P.sweep_out->pivotmin = P.st[0].pivotmin;
    for (int v = 0; v < P.P; ++v) P.st[v].pivotmax_prev = P.st[v].pivotmax;   // dmrgg.f90:961
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k_run_begin(DevPlan P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    P.ctrl->ready = 0; P.ctrl->strike = 0; P.ctrl->error = 0; P.ctrl->nsweeps = 0;
    P.ctrl->t0_ns = globaltimer_ns();
}
// record of sweep `it` (after the quadrature) + the exit test of dmrgg.f90:1010-1019
__global__ void k_sweep_log(DevPlan P, int it, int maxrank) {
    if (P.ctrl->ready) return;
    for (int x = threadIdx.x; x <= P.d; x += blockDim.x) P.rklog[(i64)it * (P.d + 1) + x] = P.rk[x];
    if (threadIdx.x != 0) return;
    SweepOut o = *P.sweep_out;
    o.t_ns = globaltimer_ns() - P.ctrl->t0_ns;
    o.valid = 1;
    P.slog[it] = o;
    P.ctrl->nsweeps = it;
    int ready = 0;
    if (maxrank > 0) ready = (it + 1 >= maxrank);
    if (P.has_accuracy) {
        if (o.pivotmax <= P.accuracy * o.amax) P.ctrl->strike += 1; else P.ctrl->strike = 0;
        ready = ready || (P.ctrl->strike >= 3);
    }
    if (P.ctrl->error) ready = 1;
    __threadfence();
    P.ctrl->ready = ready;
}

// ----------------------------------------------------------------------------
// neighbour exchange between virtual ranks (dmrgg.f90:872-958 LEFT, dmrggmp.f90:572-629 RIGHT).
// With one padded copy of every core in HBM the new column / row of the shared core is already in place;
// what remains is the corner (both sides evaluate the same n(c) points; each counts them) and the factor extensions.
// blockIdx.y = boundary b between virtual ranks b and b+1, shared core c = own[b+1].
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_exchange_corner(DevPlan P) {
    if (P.ctrl->ready) return;
    extern __shared__ double smem[];
    __shared__ Partial shp[32];
    const int b = blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    if (!(rc1 > rc1s && rc > rcs)) return;
    const double* A = stage_aux<KIND>(P, smem);
    const int nc = P.n[c];
    double* argc = P.arg + P.coreOff[c];
    Partial best = amax_init();
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        PointSrc s;
        s.L = P.Lidx + P.offL[c - 1]; s.nl = c - 1; s.i = rc1; s.j = j + 1; s.k = 0; s.hask = 0;
        s.R = P.Ridx + P.offR[c]; s.q = rc; s.Rmax = P.Rmax;
        double f = eval_point<KIND>(P, s, A);
        argc[(rc1 - 1) + (i64)P.Rmax * (j + (i64)nc * (rc - 1))] = f;
        amax_take(best, f, j);
    }
    best = amax_block(best, shp);
    if (threadIdx.x == 0) {
        // virtual rank b+1 is also touched by the CTA of boundary b+1: atomics (amax >= 0, so the bit pattern orders like the value)
        for (int v = b; v <= b + 1; ++v) {
            atomicMax((long long*)&P.st[v].amax, __double_as_longlong(best.absv));
            atomicAdd((unsigned long long*)&P.st[v].neval, (unsigned long long)nc);
        }
    }
}
__global__ void k_exchange_extend(DevPlan P) {
    if (P.ctrl->ready) return;
    const int b = blockIdx.y;
    const int c = P.own[b + 1];
    const int rc1 = P.rk[c - 1], rc1s = P.rks[c - 1], rc = P.rk[c], rcs = P.rks[c];
    const int nc = P.n[c];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const double* argc = P.arg + P.coreOff[c];
    if (e < nc) {
        if (rc > rcs) {
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
        if (rc1 > rc1s) {
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
    const int p = blockIdx.y + 1;
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
    if (P.ctrl->ready) return;
    const int p = blockIdx.x + 1;
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
    const int v = blockIdx.x;
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
__global__ void k_lua_r(DevPlan P) {   // d2_luar(n*r1, r0, inv(p-1)): thread per column (j,k)
    const int p = blockIdx.y + 1;
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
__global__ void k_lua_l(DevPlan P) {   // d2_lual(r0*n, r1, inv(p)): thread per row (i,j); cores 1..d-1
    const int p = blockIdx.y + 1;
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
    const int r0 = P.rk[p - 1], r1 = P.rk[p], n = P.n[p];
    const double* a = P.arg + P.coreOff[p];
    const i64 tot = (i64)r0 * n * r1;
    for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        int i = (int)(e % r0); i64 jk = e / r0;
        out[e] = a[i + (i64)P.Rmax * jk];
    }
}

// ----------------------------------------------------------------------------
// initial cross (dmrgg.f90:150-232)
// ----------------------------------------------------------------------------
template <int KIND>
__global__ void k_init_search(DevPlan P, int nn, int snum, double* b) {
    extern __shared__ double smem[];
    const double* A = stage_aux<KIND>(P, smem);
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < nn * snum; x += gridDim.x * blockDim.x) {
        DiagSrc s; s.n = P.n; s.k = x % nn + 1; s.s = x / nn;
        b[x] = eval_point<KIND>(P, s, A);
    }
}
// fiber of core p through the initial cross: arg(p)(1,j,1) = f(ind0 with position p := j); tables hold pivot 1 already
template <int KIND>
__global__ void k_init_cross(DevPlan P) {
    extern __shared__ double smem[];
    const int p = blockIdx.y + 1;
    const double* A = stage_aux<KIND>(P, smem);
    const int n = P.n[p];
    double* a = P.arg + P.coreOff[p];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        PointSrc s;
        s.L = P.Lidx + P.offL[p - 1]; s.nl = p - 1; s.i = 1; s.j = j + 1; s.k = 0; s.hask = 0;
        s.R = P.Ridx + P.offR[p]; s.q = 1; s.Rmax = P.Rmax;
        a[(i64)P.Rmax * j] = eval_point<KIND>(P, s, A);
    }
}

// factors of the initial cross (dmrgg.f90:234-248): inv(p)(1) = pivot, col(p) = arg(p)/pivot (d2_lual, r = 1),
// row(p) = arg(p) (d2_luar with r = 1 is the identity).  blockIdx.y = core - 1.
__global__ void k_init_factors(DevPlan P) {
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

}  // namespace ttc
