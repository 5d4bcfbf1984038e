// =============================================================================
//  ttcross_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.
//
//  A plain C++ restatement of the reference's greedy TT-cross sweep
//  (aukeschaap/ttcross, Fortran).  Only tests/, __graft_entry__.smoke() and
//  bench.py's cpu_baseline / --impl reference legs may load this file's shared
//  object.  The product (ttcross_b200/) never includes, links or calls it.
//
//  PARITY STATUS: "parity unpinned" for pivots / ranks / neval — the reference
//  ships no golden vectors for them and cannot be compiled in the build
//  container (no Fortran, MPI, BLAS).  The oracle is pinned only by the
//  reference's analytic integral values (test_crs_ising.f90:71-100,
//  test_crs_stdnorm.f90:83, test_crs_mvn.f90:83), by the one known-answer table
//  the reference ships (get_reference_val, test_crs_chf.f90:232-271: 32 complex
//  characteristic-function values of the MVN cross + ztt_quad pipeline at DIM = 4,
//  reproduced to < 8e-5 absolute = the table's own resolution, tests/test_chf.py)
//  and by brute-force tensor checks in tests/.
//
//  What is restated (reference file:line):
//    dtt_dmrgg        lib/dmrgg.f90:11-1050      (main routine)
//    dmrgg_fun        lib/dmrgg.f90:1053-1078    (index assembly)
//    dtt_lua          lib/dmrgg.f90:1169-1258
//    dtt_quad         lib/dmrgg.f90:1261-1415
//    RIGHT exchange   lib/dmrggmp.f90:572-629    (absent from dmrgg.f90)
//    d2_lual/d2_luar  lib/lr.f90:124-154
//    lottery2/find_d  lib/rnd.f90:105-144
//    share            lib/default.f90:80-97
//    lgwt             lib/quad.f90:97-131
//    dtt_rank (erank) lib/tt.f90:1228-1245
//    ort0_d           lib/ort.f90:17-81 (LAPACK dgeqrf + dorgqr restated as dgeqr2 + dorg2r; pinned against numpy's LAPACK)
//    dtt_ort, dtt_svd lib/tt.f90:130-198, 307-368 with chop of lib/mat.f90:433-455 (first row of SURVEY 8(f))
//    COS coefficient  lib/coefficients.f90:33-65, lib/funcs.f90:8-26, lib/s_vectors.f90:7-29
//    integrands       test_crs_ising.f90:176-218, test_crs_stdnorm.f90:154-170,
//                     lib/mvn_pdf.f90:63-83 + test_crs_mvn.f90:156-172
//    BLAS semantics   netlib reference order (idamax first-max, sequential ddot,
//                     dgemv 'n' column-accumulate, dgemv 't' dot-then-add,
//                     dgemm column/l-loop), no FMA.  
//
//  MPI ranks are simulated as P "virtual ranks" inside one process; every
//  message of the reference is a copy between Rank objects, delivered with
//  sendrecv semantics (all sends read pre-exchange state).
//
//  The reference never seeds random_number; here the uniform stream
//  is an explicit input: a counter-based splitmix64 stream per virtual rank.
// =============================================================================
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/ttc_detexp.h"   // the deterministic exp both sides use in parity mode (a product header; the product never includes oracle/)

namespace {
double g_visit_seconds = 0;      // time spent in the bond visits (the part the ranks run concurrently), for tto_visit_seconds
int g_rank_concurrency = 1;     // tto_set_rank_concurrency: 0 runs the virtual ranks one after another

using i64 = long long;
using u64 = unsigned long long;

// ----------------------------------------------------------------------------
// uniform stream (explicit input replacing gfortran's unseeded random_number)
// ----------------------------------------------------------------------------
inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline double stream_uniform(u64 seed, int vrank, u64 k) {
    const u64 G = 0x9E3779B97F4A7C15ULL;
    u64 base = mix64(seed + G * (u64)(vrank + 1));
    u64 z = mix64(base + G * (k + 1));
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);  // [0,1)
}

// ----------------------------------------------------------------------------
// netlib-order BLAS (SURVEY §8c "exact semantics")
// ----------------------------------------------------------------------------
int idamax(i64 n, const double* x, i64 inc = 1) {  // 1-based; first max of |x|
    if (n < 1) return 0;
    i64 best = 1;
    double dmax = std::fabs(x[0]);
    for (i64 i = 2; i <= n; ++i) {
        double v = std::fabs(x[(i - 1) * inc]);
        if (v > dmax) { best = i; dmax = v; }
    }
    return (int)best;
}
double ddot(int n, const double* x, i64 incx, const double* y, i64 incy) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t = t + x[i * incx] * y[i * incy];
    return t;
}
// y := y + alpha*A*x ('n'), beta is 1 or 0
void dgemv_n(i64 m, int n, double alpha, const double* A, i64 lda, const double* x, i64 incx,
             double beta, double* y) {
    if (beta == 0.0) for (i64 i = 0; i < m; ++i) y[i] = 0.0;
    for (int j = 0; j < n; ++j) {
        double temp = alpha * x[j * incx];
        const double* a = A + (i64)j * lda;
        for (i64 i = 0; i < m; ++i) y[i] = y[i] + temp * a[i];
    }
}
// y := y + alpha*A^T*x ('t'), A is m x n, y has n entries with stride incy
void dgemv_t(int m, i64 n, double alpha, const double* A, i64 lda, const double* x, i64 incx,
             double* y, i64 incy) {
    for (i64 j = 0; j < n; ++j) {
        double temp = 0.0;
        const double* a = A + j * lda;
        for (int i = 0; i < m; ++i) temp = temp + a[i] * x[i * incx];
        y[j * incy] = y[j * incy] + alpha * temp;
    }
}
// C := alpha*A*B + beta*C, 'n','n'
void dgemm_nn(i64 m, i64 n, int k, double alpha, const double* A, i64 lda, const double* B, i64 ldb,
              double beta, double* C, i64 ldc) {
    for (i64 j = 0; j < n; ++j) {
        double* c = C + j * ldc;
        if (beta == 0.0) for (i64 i = 0; i < m; ++i) c[i] = 0.0;
        for (int l = 0; l < k; ++l) {
            double temp = alpha * B[l + j * ldb];
            const double* a = A + (i64)l * lda;
            for (i64 i = 0; i < m; ++i) c[i] = c[i] + temp * a[i];
        }
    }
}

// lib/lr.f90:124-139
void d2_lual(i64 m, int r, const double* g, double* col, int from = 1) {
    for (int p = from; p <= r; ++p) {
        if (p > 1) dgemv_n(m, p - 1, -1.0, col, m, g + (p * p - p + 1 - 1), 1, 1.0, col + (i64)(p - 1) * m);
        double s = 1.0 / g[p * p - 1];
        double* c = col + (i64)(p - 1) * m;
        for (i64 i = 0; i < m; ++i) c[i] = s * c[i];
    }
}
// lib/lr.f90:140-154
void d2_luar(i64 n, int r, const double* g, double* row, int from = 1) {
    for (int p = from; p <= r; ++p) {
        if (p > 1) dgemv_t(p - 1, n, -1.0, row, r, g + (p * p - 2 * p + 2 - 1), 1, row + (p - 1), r);
    }
}

// lib/rnd.f90:128-144
int find_d(int n, const double* x /*1-based x(1..n) == x[0..n-1]*/, double y) {
    if (n == 0) return 0;
    if (y < x[0]) return 0;
    if (x[n - 1] <= y) return n;
    int s = 1, t = n, pos = (t + s) / 2;
    while (t - s > 1) {
        if (y < x[pos - 1]) t = pos; else s = pos;
        pos = (s + t) / 2;
    }
    return pos;
}
// lib/rnd.f90:105-126.  `d` holds 2*npnt uniforms, column-major d(npnt,2).
void lottery2(int npnt, int m, int n, const double* wcol, const double* wrow, const double* d, int* points /*npnt x 2 col-major*/) {
    std::vector<double> pcol(m + 1), prow(n + 1);
    double scol = 0.0, srow = 0.0;
    for (int i = 0; i < m; ++i) scol = scol + std::fabs(wcol[i]);
    for (int j = 0; j < n; ++j) srow = srow + std::fabs(wrow[j]);
    pcol[0] = 0.0; for (int i = 1; i <= m; ++i) pcol[i] = pcol[i - 1] + std::fabs(wcol[i - 1]) / scol;
    prow[0] = 0.0; for (int j = 1; j <= n; ++j) prow[j] = prow[j - 1] + std::fabs(wrow[j - 1]) / srow;
    for (int ip = 0; ip < npnt; ++ip) {
        int a = find_d(m + 1, pcol.data(), d[ip]);        if (a > m) a = m;
        int b = find_d(n + 1, prow.data(), d[npnt + ip]); if (b > n) b = n;
        points[ip] = a; points[npnt + ip] = b;
    }
}

// lib/default.f90:80-97
void share(int first, int last, int nproc, int* own /*0..nproc*/) {
    own[0] = first;
    for (int p = 1; p <= nproc - 1; ++p) own[p] = first + (int)((double)(last - first + 1) * (double)p / nproc);
    own[nproc] = last + 1;
}

// lib/quad.f90:97-131
void lgwt(int n, double* x, double* w) {
    const double tpi = 6.283185307179586476925286766559005768394338798750211641949889184615632812572417997256069650684234;
    double small = 5 * 2.220446049250313e-16;
    int m = (n + 1) / 2;
    for (int i = 1; i <= m; ++i) {
        double z = std::cos((tpi * (4 * i - 1)) / (8 * n + 4));
        double p1, p2, p3, pp, z1;
        do {
            p1 = 1.0; p2 = 0.0;
            for (int j = 1; j <= n; ++j) { p3 = p2; p2 = p1; p1 = ((2 * j - 1) * z * p2 - (j - 1) * p3) / j; }
            pp = n * (z * p1 - p2) / (z * z - 1);
            z1 = z;
            z = z1 - p1 / pp;
        } while (std::fabs(z - z1) > small);
        x[i - 1] = -z; x[n - i] = z;
        w[i - 1] = 2.0 / ((1 - z * z) * pp * pp);
        w[n - i] = w[i - 1];
    }
}

// ----------------------------------------------------------------------------
// integrands.  ind is 1-based values in a 0-based C array ind[0..m-1].
// ----------------------------------------------------------------------------
enum Kind { ISING = 1, STDNORM = 4, MVN = 5 };

struct Problem {
    int kind = ISING;
    int d = 0;
    std::vector<int> n;         // n[0..d-1]
    std::vector<double> par;
    std::vector<double> aux;    // MVN: mu(d) | inv_cov(d,d) column-major | denom
    int exp_mode = 0;           // 1: exp through include/ttc_detexp.h (parity mode; the product has the same switch)
};
inline double prob_exp(int mode, double x) { return mode ? ttc_det_exp(x) : std::exp(x); }

// test_crs_ising.f90:176-218
double f_ising(int m, const int* ind, const int* n, const double* par) {
    int id = (int)par[2 * n[0]];
    const double* nodes = par - 1;             // nodes[ind] == par(nodes+ind)
    const double* weights = par + n[0] - 1;
    double a = 0, b = 0, f = 0;
    if (id == 2 || id == 3) {
        a = 1.0;
        for (int i = 0; i <= m; ++i) {
            double uij = 1.0;
            for (int j = i + 1; j <= m; ++j) {
                uij = uij * nodes[ind[j - 1]];
                double t = (uij - 1.0) / (uij + 1.0);
                a = a * (t * t);
            }
        }
    }
    if (id == 1 || id == 2) {
        double v = 1.0, w = 1.0, vk = 1.0, wk = 1.0;
        for (int i = 1; i <= m; ++i) {
            vk = vk * nodes[ind[m - i]];
            wk = wk * nodes[ind[i - 1]];
            v = v + vk;
            w = w + wk;
        }
        b = 1.0 / (v * w);
    }
    switch (id) {
        case 1: f = 2 * b; break;
        case 2: f = 2 * a * b; break;
        case 3: f = 2 * a; break;
        default: std::fprintf(stderr, "unknown id: %d\n", id); std::abort();
    }
    for (int i = 1; i <= m; ++i) f = f * weights[ind[i - 1]];
    return f;
}
// test_crs_stdnorm.f90:154-170
double f_stdnorm(int m, const int* ind, const int*, const double* par, int exp_mode) {
    double s = 0.0;
    for (int i = 0; i < m; ++i) { double x = par[ind[i] - 1]; s = s + x * x; }
    return prob_exp(exp_mode, -s);
}
// lib/mvn_pdf.f90:63-83 through test_crs_mvn.f90:156-172
double f_mvn(int m, const int* ind, const int*, const double* par, const double* aux, int exp_mode) {
    const double* mu = aux; const double* A = aux + m; double denom = aux[m + (i64)m * m];
    std::vector<double> diff(m);
    for (int i = 0; i < m; ++i) diff[i] = par[ind[i] - 1] - mu[i];
    double e = 0.0;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) e = e + diff[i] * A[i + (i64)j * m] * diff[j];
    return prob_exp(exp_mode, -0.5 * e) / denom;
}
// lib/coefficients.f90:33-65 (calc_coefficient) with lib/funcs.f90:8-26 (gaussian_chf_nd) and lib/s_vectors.f90:7-29.
// aux = mu(d) | sigma(d,d) column-major | lower | upper.  x**n: libgcc __powidf2; matmul(sigma, t): column by column from 0;
// cexp(z) = exp(re z) (cos(im z), sin(im z)); complex product (a+bi)(c+di) -> real part ac - bd.
static double powi_gcc(double x, int m) {
    unsigned int n = m < 0 ? (unsigned)(-m) : (unsigned)m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n % 2) y = y * x; }
    return m < 0 ? 1.0 / y : y;
}
double f_coscoef(int m, const int* ind, const double* aux) {
    const double* mu = aux; const double* sg = aux + m;
    const double lower = aux[m + (i64)m * m], upper = aux[m + (i64)m * m + 1];
    const double pi = 3.14159265358979323846;
    const double oob = 1 / (upper - lower);
    const double factor = 2.0 * powi_gcc(oob, m);
    std::vector<double> t(m), y(m);
    double real_sum = 0.0;
    const unsigned ns = 1u << (m - 1);
    for (unsigned i = 0; i < ns; ++i) {
        for (int j = 0; j < m; ++j) {
            const int sj = (j == 0) ? 1 : (((i >> (j - 1)) & 1u) ? -1 : 1);
            t[j] = (((pi * (double)sj) * (double)(ind[j] - 1)) * oob);
        }
        double dot_mu = 0.0, st = 0.0;
        for (int j = 0; j < m; ++j) { dot_mu = dot_mu + t[j] * mu[j]; st = st + t[j]; }
        for (int a = 0; a < m; ++a) y[a] = 0.0;
        for (int b = 0; b < m; ++b) for (int a = 0; a < m; ++a) y[a] = y[a] + sg[a + (i64)b * m] * t[b];
        double quad = 0.0;
        for (int a = 0; a < m; ++a) quad = quad + y[a] * t[a];
        const double E = std::exp(-0.5 * quad);
        const double c2 = std::cos(dot_mu), s2 = std::sin(dot_mu);
        const double c1 = std::cos(-lower * st), s1 = std::sin(-lower * st);
        real_sum = real_sum + (c1 * (E * c2) - s1 * (E * s2));
    }
    return factor * real_sum;
}
double fun(const Problem& P, const int* ind) {
    switch (P.kind) {
        case 6: return f_coscoef(P.d, ind, P.aux.data());
        case ISING: return f_ising(P.d, ind, P.n.data(), P.par.data());
        case STDNORM: return f_stdnorm(P.d, ind, P.n.data(), P.par.data(), P.exp_mode);
        case MVN: return f_mvn(P.d, ind, P.n.data(), P.par.data(), P.aux.data(), P.exp_mode);
    }
    std::abort();
}

// ----------------------------------------------------------------------------
// TT container: core(i,j,k), 1-based, column-major packed (tt.f90:18-26)
// ----------------------------------------------------------------------------
struct Core {
    int r0 = 0, n = 0, r1 = 0;
    std::vector<double> a;
    void alloc(int r0_, int n_, int r1_) { r0 = r0_; n = n_; r1 = r1_; a.assign((size_t)r0 * n * r1, 0.0); }
    double& at(int i, int j, int k) { return a[(size_t)(i - 1) + (size_t)r0 * ((j - 1) + (size_t)n * (k - 1))]; }
    const double& at(int i, int j, int k) const { return a[(size_t)(i - 1) + (size_t)r0 * ((j - 1) + (size_t)n * (k - 1))]; }
    double* p() { return a.data(); }
};

using Vip = std::vector<std::array<int, 4>>;  // vip(p)%p(1:4, s) == vip[s-1][0..3]

struct PivRec { int it, vrank, bond, ii, jj, kk, qq, upd; double pivot; };

struct Rank {
    int me = 0;
    std::vector<int> r, rr;              // index 0..d
    std::vector<Vip> vip;                // index 0..d
    std::vector<std::vector<double>> inv;// index 0..d
    std::vector<Core> arg, col, row;     // index 1..d (0 unused)
    std::vector<std::array<int, 4>> tape, tmpp;  // index 0..d
    std::vector<char> upd;               // index 0..d
    double amax = 0, pivotmax = -1, pivotmin = -1, pivotmax_prev = 0;
    i64 nevalloc = 0;
    u64 rng_k = 0;
};

struct Result {
    std::vector<int> ranks;
    i64 neval = 0;
    std::vector<double> vals;      // per sweep, index 0 = initial
    std::vector<i64> nevals;       // per sweep (rank-0 reduce)
    std::vector<double> amaxs, pivotmaxs, eranks, times;
    std::vector<PivRec> pivlog;
    std::vector<Core> cores;       // finalised TT, index 1..d (as held by owners)
    std::string text;
    double quad_final = 0;
    double seconds = 0;
    int nsweeps = 0;
};

struct Oracle {
    Problem prob;
    Result res;
};

// dmrgg.f90:1053-1078
double dmrgg_fun(const Problem& P, int i, int j, int k, int q, int p, int l, int m, const std::vector<Vip>& vip) {
    int ind[2048];
    int t = i;
    for (int s = p - 1; s >= l; --s) {
        ind[s - l] = vip[s][t - 1][1];
        t = vip[s][t - 1][0];
    }
    ind[p - 1] = j;      // ind(p)
    ind[p] = k;          // ind(p+1)
    t = q;
    for (int s = p + 1; s <= m - 1; ++s) {
        ind[s + 1 - l] = vip[s][t - 1][2];
        t = vip[s][t - 1][3];
    }
    return fun(P, ind);
}

// tt.f90:1228-1245
double erank(int l, int m, const std::vector<int>& n /*1-based via n[k-1]*/, const std::vector<int>& r) {
    int d = m - l + 1;
    if (d <= 0) return -1.0;
    if (d == 1) return 0.0;
    double rk = 0.0;
    for (int i = l; i <= m; ++i) rk = rk + (double)(r[i - 1] * n[i - 1] * r[i]);
    if (rk == 0.0) return rk;
    int b = r[l - 1] * n[l - 1] + n[m - 1] * r[m];
    if (d == 2) return rk / b;
    int a = 0;
    for (int i = l + 1; i <= m - 1; ++i) a += n[i - 1];
    return (std::sqrt((double)b * b + 4.0 * a * rk) - b) / (2.0 * a);
}

// Fortran-style 'e' edit descriptor (0.dddE+xx), e.g. e20.14, e9.3, e8.3
std::string fmt_e(double v, int w, int dgt) {
    char buf[128];
    if (v == 0.0 || !std::isfinite(v)) {
        std::string s = (v == 0.0) ? "0." + std::string(dgt, '0') + "E+00" : (std::isnan(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity"));
        if ((int)s.size() < w) s = std::string(w - s.size(), ' ') + s;
        return s;
    }
    // exact decimal conversion with dgt significant digits: d.ddd e±xx  ->  0.dddd E±(xx+1)
    std::snprintf(buf, sizeof buf, "%.*e", dgt - 1, std::fabs(v));
    std::string t = buf;
    size_t epos = t.find('e');
    int ex = std::atoi(t.c_str() + epos + 1) + 1;
    std::string digits; for (size_t i = 0; i < epos; ++i) if (t[i] != '.') digits += t[i];
    std::string s = std::string(v < 0 ? "-" : "") + "0." + digits;
    char eb[16]; std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
    s += eb;
    if ((int)s.size() > w && s.size() >= 2 && s[s[0] == '-' ? 1 : 0] == '0') s.erase(s[0] == '-' ? 1 : 0, 1);  // drop leading 0 like gfortran
    if ((int)s.size() > w) s = std::string(w, '*');
    if ((int)s.size() < w) s = std::string(w - s.size(), ' ') + s;
    return s;
}

// ----------------------------------------------------------------------------
// dtt_lua over virtual ranks (dmrgg.f90:1169-1258).  `cores[v]` is rank v's TT.
// ----------------------------------------------------------------------------
void dtt_lua_all(int P, const std::vector<int>& own, std::vector<std::vector<Core>*>& cores,
                 std::vector<Rank>& R) {
    // inv hand-off to the right neighbour (sendrecv semantics)
    if (P > 1) {
        std::vector<std::vector<double>> msg(P);
        for (int me = 0; me < P - 1; ++me) { int q = own[me + 1] - 1; msg[me] = R[me].inv[q]; }
        for (int me = 1; me < P; ++me) {
            int p = own[me] - 1;
            int rp = R[me].r[p];
            if ((int)msg[me - 1].size() != rp * rp) {
                std::fprintf(stderr, "oracle: dtt_lua inv size mismatch at rank %d bond %d: %zu vs %d\n", me, p, msg[me - 1].size(), rp * rp);
                std::abort();
            }
            R[me].inv[p] = msg[me - 1];
        }
    }
    for (int me = 0; me < P; ++me) {
        std::vector<Core>& c = *cores[me];
        const std::vector<int>& r = R[me].r;
        for (int p = own[me]; p <= own[me + 1] - 1; ++p) {
            d2_luar((i64)c[p].n * c[p].r1, c[p].r0, R[me].inv[p - 1].data(), c[p].p());
            d2_lual((i64)c[p].r0 * c[p].n, c[p].r1, R[me].inv[p].data(), c[p].p());
            (void)r;
        }
        if (me == P - 1) {
            int m = own[me + 1];
            d2_luar((i64)c[m].n * c[m].r1, c[m].r0, R[me].inv[m - 1].data(), c[m].p());
        }
    }
}

// dmrgg.f90:1261-1415.  quad: per-core weight vectors (index 1..d) or empty.
double dtt_quad_all(int P, const std::vector<int>& own, int m, std::vector<std::vector<Core>*>& cores,
                    const std::vector<std::vector<double>>* quad) {
    struct Mat { int m = 0, n = 0; std::vector<double> a; };
    std::vector<Mat> prev(P);
    for (int me = 0; me < P; ++me) {
        std::vector<Core>& c = *cores[me];
        int first = own[me], last = own[me + 1] - 1;
        if (me == P - 1) last = m;
        Mat pv;
        for (int p = first; p <= last; ++p) {
            Mat curr; curr.m = c[p].r0; curr.n = c[p].r1; curr.a.assign((size_t)curr.m * curr.n, 0.0);
            if (quad) {
                for (int k = 1; k <= c[p].r1; ++k)
                    dgemv_n(c[p].r0, c[p].n, 1.0, &c[p].at(1, 1, k), c[p].r0, (*quad)[p].data(), 1, 0.0, &curr.a[(size_t)(k - 1) * curr.m]);
            } else {
                for (int k = 1; k <= c[p].r1; ++k)
                    for (int i = 1; i <= c[p].r0; ++i) {
                        double s = 0.0;
                        for (int j = 1; j <= c[p].n; ++j) s = s + c[p].at(i, j, k);
                        curr.a[(size_t)(i - 1) + (size_t)(k - 1) * curr.m] = s;
                    }
            }
            if (p == first) pv = curr;
            else {
                Mat next; next.m = pv.m; next.n = curr.n; next.a.assign((size_t)next.m * next.n, 0.0);
                dgemm_nn(pv.m, curr.n, curr.m, 1.0, pv.a.data(), pv.m, curr.a.data(), curr.m, 0.0, next.a.data(), next.m);
                pv = next;
            }
        }
        prev[me] = pv;
    }
    // binary tree reduction over ranks (dmrgg.f90:1355-1405)
    for (int q = 1; q < P; q *= 2) {
        for (int me = 0; me < P; ++me) {
            if (me % (2 * q) == 0) {
                int her = me + q;
                if (her < P) {
                    Mat& a = prev[me]; Mat& b = prev[her];
                    if (a.n != b.m) { std::fprintf(stderr, "oracle: dtt_quad size mismatch %d %d | %d %d\n", a.m, a.n, b.m, b.n); std::abort(); }
                    Mat next; next.m = a.m; next.n = b.n; next.a.assign((size_t)next.m * next.n, 0.0);
                    dgemm_nn(a.m, b.n, a.n, 1.0, a.a.data(), a.m, b.a.data(), b.m, 0.0, next.a.data(), next.m);
                    a = next;
                }
            }
        }
    }
    return prev[0].a[0];
}

// ----------------------------------------------------------------------------
// dtt_dmrgg (dmrgg.f90:11-1050), P virtual ranks
// ----------------------------------------------------------------------------
struct RunArgs {
    int maxrank = -1;          // <=0: absent
    double accuracy = -1;      // <0: absent
    int piv = 3;
    int P = 1;
    std::vector<int> own;      // empty: share()
    std::vector<std::vector<double>> quad;  // index 1..d, empty: absent
    bool has_tru = false; double tru = 0;
    u64 seed = 1;
    int verbose = 0;
};

int run(Oracle& O, const RunArgs& A) {
    const Problem& prob = O.prob;
    Result& res = O.res; res = Result();
    auto tstart = std::chrono::steady_clock::now();
    auto timef = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - tstart).count(); };
    const int l = 1, m = prob.d, nproc = A.P;
    const std::vector<int>& n = prob.n;  // n(p) == n[p-1]
    auto N = [&](int p) { return n[p - 1]; };
    const int piv = A.piv;
    const double eps = 2.220446049250313e-16;
    const double small_element = 10 * eps, small_pivot = 1.e-5;
    const bool has_quad = !A.quad.empty();
    char line[512];

    if (nproc >= m) { res.text += "nproc exceeds or equal dimension, cannot proceed\n"; return 1; }
    if (piv < -1) { res.text += "dtt_dmrgg: unknown pivoting\n"; return 2; }
    std::vector<int> own(nproc + 1);
    if (!A.own.empty()) own = A.own; else share(l, m - 1, nproc, own.data());

    std::vector<Rank> R(nproc);
    for (int me = 0; me < nproc; ++me) {
        Rank& K = R[me]; K.me = me;
        K.r.assign(m + 1, 1); K.rr.assign(m + 1, 1);
        K.vip.assign(m + 1, Vip(1, {1, 1, 1, 1}));
        K.inv.assign(m + 1, std::vector<double>(1, 1.0));
        K.arg.resize(m + 1); K.col.resize(m + 1); K.row.resize(m + 1);
        for (int p = l; p <= m; ++p) K.arg[p].alloc(1, N(p), 1);
        K.tape.assign(m + 1, {-1, -1, -1, -1}); K.tmpp.assign(m + 1, {-2, -2, -2, -2});
        K.upd.assign(m + 1, 0);
    }

    // ---- locating initial cross (dmrgg.f90:150-217)
    const int smin = 8;
    int snum = std::max(smin, nproc);
    std::vector<int> shifts(nproc + 1);
    for (int p = 0; p < nproc; ++p) shifts[p] = (int)((double)snum * (double)p / nproc);
    shifts[nproc] = snum;
    int nn = *std::min_element(n.begin(), n.end());
    std::vector<double> bbv(nproc); std::vector<int> bbi(nproc);
    for (int me = 0; me < nproc; ++me) {
        int ihave = shifts[me + 1] - shifts[me];
        int nlot = nn * ihave;
        std::vector<double> b(nlot);
        for (int s = 0; s < ihave; ++s) {
#pragma omp parallel for schedule(static)
            for (int k = 1; k <= nn; ++k) {
                int ind[2048];
                for (int p = l; p <= m; ++p) ind[p - 1] = (k - 1 + (s + shifts[me]) * (p - 1)) % N(p) + 1;
                b[k - 1 + s * nn] = fun(prob, ind);
            }
        }
        int ilot = idamax(nlot, b.data());
        R[me].amax = std::fabs(b[ilot - 1]);
        R[me].nevalloc = nlot;
        ilot = ilot + nn * shifts[me];
        bbv[me] = R[me].amax; bbi[me] = ilot;
    }
    // MPI_MAXLOC: max value, lowest index among ties
    double gmax = bbv[0]; int gilot = bbi[0];
    for (int me = 1; me < nproc; ++me) {
        if (bbv[me] > gmax || (bbv[me] == gmax && bbi[me] < gilot)) { gmax = bbv[me]; gilot = bbi[me]; }
    }
    std::vector<int> ind0(m + 2, 1);
    {
        int s = (gilot - 1) / nn, k = (gilot - 1) % nn + 1;
        for (int p = l; p <= m; ++p) ind0[p] = (k - 1 + s * (p - 1)) % N(p) + 1;
    }
    for (int me = 0; me < nproc; ++me) {
        Rank& K = R[me];
        if (nproc > 1) K.amax = gmax;
        K.vip[l - 1][0] = {1, 1, 1, 1};
        for (int p = l; p <= m - 1; ++p) K.vip[p][0] = {1, ind0[p], ind0[p + 1], 1};
        K.vip[m][0] = {1, 1, 1, 1};
    }

    // ---- initial cross (dmrgg.f90:220-248)
    for (int me = 0; me < nproc; ++me) {
        Rank& K = R[me];
        for (int p = own[me]; p <= own[me + 1]; ++p) {
#pragma omp parallel for schedule(static)
            for (int j = 1; j <= N(p); ++j)
                K.arg[p].at(1, j, 1) = dmrgg_fun(prob, 1, j, (p + 1 <= m ? ind0[p + 1] : 1), 1, p, l, m, K.vip);
            K.nevalloc += N(p);
            for (int j = 1; j <= N(p); ++j) K.amax = std::max(K.amax, std::fabs(K.arg[p].at(1, j, 1)));
        }
        K.pivotmax_prev = K.amax;
        for (int p = own[me]; p <= own[me + 1] - 1; ++p) K.inv[p][0] = K.arg[p].at(1, ind0[p], 1);
        K.col = K.arg; K.row = K.arg;
        for (int p = own[me]; p <= own[me + 1] - 1; ++p) {
            d2_lual(N(p), K.r[p], K.inv[p].data(), K.col[p].p());
            d2_luar(N(p), K.r[p], K.inv[p].data(), K.row[p + 1].p());
        }
    }
    // NOTE (dmrgg.f90:224): for p == m the reference reads ind(p+1) = ind(m+1), which the forall at :207
    // never wrote; dmrgg_fun then overwrites position m+1 of a local scratch and the integrand never
    // reads it, so any value is equivalent.  We pass 1.

    double val = 0, val_prev = 0;
    if (has_quad) {
        std::vector<double> part(nproc);
        for (int me = 0; me < nproc; ++me) {
            Rank& K = R[me];
            double v = 1.0;
            for (int p = own[me]; p <= own[me + 1] - 1; ++p)
                v = v * ddot(N(p), K.arg[p].p(), 1, A.quad[p].data(), 1) / K.inv[p][0];
            if (me == nproc - 1) v = v * ddot(N(m), K.arg[m].p(), 1, A.quad[m].data(), 1);
            part[me] = v;
        }
        val = part[0];
        for (int me = 1; me < nproc; ++me) val = val * part[me];  // MPI_PROD in rank order (our definition)
        val_prev = val;
    }
    i64 nevalall = 0; for (auto& K : R) nevalall += K.nevalloc;

    std::vector<int> rrep(m + 1);
    auto report_erank = [&]() { return erank(l, m, n, R[0].r); };
    {
        double t2 = timef();
        std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", 0, "::", report_erank(), fmt_e(t2, 9, 3).c_str(), nevalall);
        std::string s = line;
        if (has_quad) s += " val " + fmt_e(val, 20, 14);
        res.text += s + "\n";
        if (A.verbose) std::puts(s.c_str());
        res.vals.push_back(val); res.nevals.push_back(nevalall); res.amaxs.push_back(R[0].amax);
        res.pivotmaxs.push_back(-1); res.eranks.push_back(report_erank()); res.times.push_back(t2);
    }

    // thread budget: ranks side by side, the rest of the threads inside each rank (nested teams)
    int outer_nt = 1, inner_nt = 1;
#ifdef _OPENMP
    {
        const int T = omp_get_max_threads();
        outer_nt = g_rank_concurrency ? std::max(1, std::min(nproc, T)) : 1;
        inner_nt = std::max(1, T / outer_nt);
        if (outer_nt > 1 && inner_nt > 1) omp_set_max_active_levels(2);
    }
#endif
    // ---- main loop (dmrgg.f90:309-1020)
    int it = 0, strike = 0;
    bool ready = false;
    if (A.maxrank > 0) ready = (it + 1 >= A.maxrank);

    while (!ready) {
        it += 1;
        int dir = 2 - it % 2;
        const char* sdir = dir == 1 ? ">>" : "<<";

        // The virtual ranks of a sweep are independent (Jacobi: each works on its own Rank object with the index sets of the
        // sweep start, dmrgg.f90:325-331) and the reference runs them as concurrent MPI ranks, each with its own OpenMP team
        // (README.md:20-21).  Same here: `outer` ranks at a time, `inner_nt` threads for a rank's evaluation loops.  The pivot
        // records are merged in rank order afterwards, so the result does not depend on the thread count.
        std::vector<std::vector<PivRec>> plog(nproc);
        const auto tv0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer_nt) if (outer_nt > 1)
        for (int me = 0; me < nproc; ++me) {
            Rank& K = R[me];
            std::vector<int>& r = K.r;
            K.rr = K.r;
            K.pivotmax = -1.0; K.pivotmin = -1.0;

            for (int pp = 1; pp <= own[me + 1] - own[me]; ++pp) {
                int p = own[me] + pp - 1;
                if (dir == 2) p = own[me + 1] - pp;
                const int r0 = r[p - 1], r1 = r[p], r2 = r[p + 1], n1 = N(p), n2 = N(p + 1);
                std::vector<double> acol1((size_t)r0 * n1), arow1((size_t)n2 * r2);
                int ii = 0, jj = 0, kk = 0, qq = 0; double pivot = 0;
                Core& colp = K.col[p]; Core& rowp1 = K.row[p + 1];
                const i64 ldc = (i64)r0 * n1;   // stride of s in col%u(p)(i,j,s)

                if (piv == -1) {
                    // full pivoting (dmrgg.f90:341-408)
                    const i64 tot = (i64)r0 * n1 * n2 * r2;
                    std::vector<double> a(tot), b(tot);
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                    for (i64 x = 1; x <= tot; ++x) {
                        i64 i = x - 1;
                        i64 q = i / ((i64)r0 * n1 * n2); i = i % ((i64)r0 * n1 * n2);
                        i64 k = i / ((i64)r0 * n1);      i = i % ((i64)r0 * n1);
                        i64 j = i / r0;                  i = i % r0;
                        a[x - 1] = dmrgg_fun(prob, (int)i + 1, (int)j + 1, (int)k + 1, (int)q + 1, p, l, m, K.vip);
                    }
                    K.nevalloc += tot;
                    i64 x = idamax(tot, a.data()) - 1;
                    K.amax = std::max(K.amax, std::fabs(a[x]));
                    b = a;
                    dgemm_nn((i64)r0 * n1, (i64)n2 * r2, r1, -1.0, colp.p(), (i64)r0 * n1, rowp1.p(), r1, 1.0, b.data(), (i64)r0 * n1);
                    x = idamax(tot, b.data()) - 1;
                    i64 y = x;
                    qq = (int)(y / ((i64)r0 * n1 * n2)) + 1; y = y % ((i64)r0 * n1 * n2);
                    kk = (int)(y / ((i64)r0 * n1)) + 1;      y = y % ((i64)r0 * n1);
                    jj = (int)(y / r0) + 1;
                    ii = (int)(y % r0) + 1;
                    pivot = b[x];
                    for (int j = 1; j <= n1; ++j) for (int i = 1; i <= r0; ++i)
                        acol1[(i - 1) + (size_t)r0 * (j - 1)] = a[(i - 1) + (size_t)r0 * ((j - 1) + (size_t)n1 * ((kk - 1) + (size_t)n2 * (qq - 1)))];
                    for (int q = 1; q <= r2; ++q) for (int k = 1; k <= n2; ++k)
                        arow1[(k - 1) + (size_t)n2 * (q - 1)] = a[(ii - 1) + (size_t)r0 * ((jj - 1) + (size_t)n1 * ((k - 1) + (size_t)n2 * (q - 1)))];
                } else {
                    // partial pivoting (dmrgg.f90:410-589)
                    const int nlot = r0 + n1 + n2 + r2;
                    std::vector<double> bcol1((size_t)r0 * n1, 1.0), brow1((size_t)n2 * r2, 1.0), b(nlot);
                    for (int s = 1; s <= r1; ++s) {
                        const auto& v = K.vip[p][s - 1];
                        bcol1[(v[0] - 1) + (size_t)r0 * (v[1] - 1)] = 0.0;
                        brow1[(v[2] - 1) + (size_t)n2 * (v[3] - 1)] = 0.0;
                    }
                    std::vector<double> u(2 * (size_t)nlot);
                    for (int x = 0; x < 2 * nlot; ++x) u[x] = stream_uniform(A.seed, me, K.rng_k + x);
                    K.rng_k += 2 * (u64)nlot;
                    std::vector<int> pts(2 * (size_t)nlot), lot(4 * (size_t)nlot);
                    lottery2(nlot, r0 * n1, n2 * r2, bcol1.data(), brow1.data(), u.data(), pts.data());
                    for (int x = 0; x < nlot; ++x) {
                        int c = pts[x], w = pts[nlot + x];
                        lot[x] = (c - 1) % r0 + 1;             // i
                        lot[nlot + x] = (c - 1) / r0 + 1;      // j
                        lot[2 * nlot + x] = (w - 1) % n2 + 1;  // k
                        lot[3 * nlot + x] = (w - 1) / n2 + 1;  // q
                    }
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                    for (int x = 0; x < nlot; ++x)
                        b[x] = dmrgg_fun(prob, lot[x], lot[nlot + x], lot[2 * nlot + x], lot[3 * nlot + x], p, l, m, K.vip);
                    K.nevalloc += nlot;
                    int ilot = idamax(nlot, b.data());
                    K.amax = std::max(K.amax, std::fabs(b[ilot - 1]));
                    for (int x = 0; x < nlot; ++x) {
                        int i = lot[x], j = lot[nlot + x], k = lot[2 * nlot + x], q = lot[3 * nlot + x];
                        b[x] = b[x] - ddot(r1, &colp.at(i, j, 1), ldc, &rowp1.at(1, k, q), 1);
                    }
                    ilot = idamax(nlot, b.data());
                    ii = lot[ilot - 1]; jj = lot[nlot + ilot - 1]; kk = lot[2 * nlot + ilot - 1]; qq = lot[3 * nlot + ilot - 1];
                    pivot = b[ilot - 1];

                    bool done = false, havecol = false, haverow = false;
                    if (piv == 0) {
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                        for (int ij = 0; ij < r0 * n1; ++ij) {
                            int j = ij / r0 + 1, i = ij % r0 + 1;
                            acol1[ij] = dmrgg_fun(prob, i, j, kk, qq, p, l, m, K.vip);
                        }
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                        for (int kq = 0; kq < n2 * r2; ++kq) {
                            int q = kq / n2 + 1, k = kq % n2 + 1;
                            arow1[kq] = dmrgg_fun(prob, ii, jj, k, q, p, l, m, K.vip);
                        }
                        K.nevalloc += (i64)r0 * n1 + (i64)n2 * r2;
                        done = havecol = haverow = true;
                    }
                    int crs = 0;
                    bool skipcol = (dir == 2);
                    while (!done) {
                        if (!skipcol) {
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                            for (int ij = 0; ij < r0 * n1; ++ij) {
                                int j = ij / r0 + 1, i = ij % r0 + 1;
                                acol1[ij] = dmrgg_fun(prob, i, j, kk, qq, p, l, m, K.vip);
                            }
                            K.nevalloc += (i64)r0 * n1;
                            int ij = idamax((i64)r0 * n1, acol1.data()) - 1;
                            K.amax = std::max(K.amax, std::fabs(acol1[ij]));
                            havecol = true; crs += 1;
                            done = havecol && haverow && (crs >= 2 * piv);
                            if (!done) {
                                bcol1 = acol1;
                                dgemv_n((i64)r0 * n1, r1, -1.0, colp.p(), (i64)r0 * n1, &rowp1.at(1, kk, qq), 1, 1.0, bcol1.data());
                                ij = idamax((i64)r0 * n1, bcol1.data()) - 1;
                                int j = ij / r0 + 1, i = ij % r0 + 1;
                                done = havecol && haverow && (i == ii && j == jj);
                                ii = i; jj = j;
                                pivot = bcol1[(ii - 1) + (size_t)r0 * (jj - 1)];
                            }
                        }
                        skipcol = false;
                        if (!done) {
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                            for (int kq = 0; kq < n2 * r2; ++kq) {
                                int q = kq / n2 + 1, k = kq % n2 + 1;
                                arow1[kq] = dmrgg_fun(prob, ii, jj, k, q, p, l, m, K.vip);
                            }
                            K.nevalloc += (i64)n2 * r2;
                            int kq = idamax((i64)n2 * r2, arow1.data()) - 1;
                            K.amax = std::max(K.amax, std::fabs(arow1[kq]));
                            haverow = true; crs += 1;
                            done = havecol && haverow && (crs >= 2 * piv);
                            if (!done) {
                                brow1 = arow1;
                                dgemv_t(r1, (i64)n2 * r2, -1.0, rowp1.p(), r1, &colp.at(ii, jj, 1), ldc, brow1.data(), 1);
                                kq = idamax((i64)n2 * r2, brow1.data()) - 1;
                                int q = kq / n2 + 1, k = kq % n2 + 1;
                                done = havecol && haverow && (k == kk && q == qq);
                                qq = q; kk = k;
                                pivot = brow1[(kk - 1) + (size_t)n2 * (qq - 1)];
                            }
                        }
                    }
                }

                // accept test + update (dmrgg.f90:598-758)
                K.tape[p] = {-1, -1, -1, -1};
                K.upd[p] = (std::fabs(pivot) > small_element * K.amax) && (std::fabs(pivot) > small_pivot * K.pivotmax_prev);
                plog[me].push_back({it, me, p, ii, jj, kk, qq, (int)K.upd[p], pivot});

                if (K.upd[p]) {
                    K.tape[p] = {ii, jj, kk, qq};
                    K.vip[p].push_back({ii, jj, kk, qq});
                    K.pivotmax = (K.pivotmax < 0.0) ? std::fabs(pivot) : std::max(K.pivotmax, std::fabs(pivot));
                    K.pivotmin = (K.pivotmin < 0.0) ? std::fabs(pivot) : std::min(K.pivotmin, std::fabs(pivot));

                    // inverse (dmrgg.f90:650-660)
                    std::vector<double>& g = K.inv[p];
                    g.resize((size_t)(r1 + 1) * (r1 + 1));
                    for (int s = 1; s <= r1; ++s) g[(size_t)r1 * r1 + s - 1] = colp.at(ii, jj, s);
                    for (int s = 1; s <= r1; ++s) g[(size_t)r1 * r1 + r1 + s - 1] = rowp1.at(s, kk, qq);
                    g[(size_t)(r1 + 1) * (r1 + 1) - 1] = pivot;

                    // arg blocks (dmrgg.f90:663-685)
                    {
                        Core nb; nb.alloc(r0, n1, r1 + 1);
                        std::copy(K.arg[p].a.begin(), K.arg[p].a.end(), nb.a.begin());
                        std::copy(acol1.begin(), acol1.end(), nb.a.begin() + (size_t)r0 * n1 * r1);
                        K.arg[p] = nb;
                        Core nr; nr.alloc(r1 + 1, n2, r2);
                        for (int q = 1; q <= r2; ++q) for (int k = 1; k <= n2; ++k) {
                            for (int s = 1; s <= r1; ++s) nr.at(s, k, q) = K.arg[p + 1].at(s, k, q);
                            nr.at(r1 + 1, k, q) = arow1[(k - 1) + (size_t)n2 * (q - 1)];
                        }
                        K.arg[p + 1] = nr;
                    }
                    // factors (dmrgg.f90:688-713)
                    {
                        Core nb; nb.alloc(r0, n1, r1 + 1);
                        std::copy(colp.a.begin(), colp.a.end(), nb.a.begin());
                        std::copy(acol1.begin(), acol1.end(), nb.a.begin() + (size_t)r0 * n1 * r1);
                        Core nr; nr.alloc(r1 + 1, n2, r2);
                        for (int q = 1; q <= r2; ++q) for (int k = 1; k <= n2; ++k) {
                            for (int s = 1; s <= r1; ++s) nr.at(s, k, q) = rowp1.at(s, k, q);
                            nr.at(r1 + 1, k, q) = arow1[(k - 1) + (size_t)n2 * (q - 1)];
                        }
                        d2_lual((i64)r0 * n1, r1 + 1, g.data(), nb.p(), r1 + 1);
                        d2_luar((i64)n2 * r2, r1 + 1, g.data(), nr.p(), r1 + 1);
                        K.col[p] = nb; K.row[p + 1] = nr;
                    }
                    if (p > own[me]) {  // left rows (dmrgg.f90:715-728)
                        Core nb; nb.alloc(r0, n1, r1 + 1);
                        std::copy(K.row[p].a.begin(), K.row[p].a.end(), nb.a.begin());
                        std::vector<double> t1(K.arg[p].a.begin() + (size_t)r0 * n1 * r1, K.arg[p].a.begin() + (size_t)r0 * n1 * (r1 + 1));
                        d2_luar(n1, r0, K.inv[p - 1].data(), t1.data());
                        std::copy(t1.begin(), t1.end(), nb.a.begin() + (size_t)r0 * n1 * r1);
                        K.row[p] = nb;
                    }
                    if (p < own[me + 1] - 1) {  // right cols (dmrgg.f90:730-749)
                        Core nr; nr.alloc(r1 + 1, n2, r2);
                        std::vector<double> t1((size_t)n2 * r2);
                        for (int q = 1; q <= r2; ++q) for (int k = 1; k <= n2; ++k) {
                            for (int s = 1; s <= r1; ++s) nr.at(s, k, q) = K.col[p + 1].at(s, k, q);
                            t1[(k - 1) + (size_t)n2 * (q - 1)] = K.arg[p + 1].at(r1 + 1, k, q);
                        }
                        d2_lual(n2, r2, K.inv[p + 1].data(), t1.data());
                        for (int q = 1; q <= r2; ++q) for (int k = 1; k <= n2; ++k) nr.at(r1 + 1, k, q) = t1[(k - 1) + (size_t)n2 * (q - 1)];
                        K.col[p + 1] = nr;
                    }
                    r[p] = r1 + 1;
                }
            }  // own bonds
        }      // ranks
        for (int me = 0; me < nproc; ++me) res.pivlog.insert(res.pivlog.end(), plog[me].begin(), plog[me].end());
        g_visit_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - tv0).count();

        if (nproc > 1) {
            // ---- tape propagation (dmrgg.f90:763-850), sendrecv semantics
            for (auto& K : R) for (int p = l - 1; p <= m; ++p) K.tmpp[p] = {-2, -2, -2, -2};
            for (int me = 0; me < nproc - 1; ++me) {         // going right: me -> me+1
                int ihave = own[me + 1] - l;
                for (int x = 0; x < ihave; ++x) R[me + 1].tmpp[l + x] = R[me].tape[l + x];
            }
            for (int me = 1; me < nproc; ++me) {             // going left: me -> me-1
                int ihave = m - own[me], q = own[me];
                for (int x = 0; x < ihave; ++x) R[me - 1].tmpp[q + x] = R[me].tape[q + x];
            }
            for (int me = 0; me < nproc; ++me) {
                Rank& K = R[me];
                for (int p = l - 1; p <= m; ++p) K.tape[p] = K.tmpp[p];
                for (int p = l; p <= m - 1; ++p) {
                    if (!(own[me] <= p && p <= own[me + 1] - 1)) {
                        K.upd[p] = (K.tape[p][0] > 0);
                        if (K.upd[p]) { K.vip[p].push_back(K.tape[p]); K.r[p] += 1; }
                    }
                }
            }
            // ---- allreduce MAX (dmrgg.f90:852-870)
            double c1 = R[0].amax, c2 = R[0].pivotmax, c3 = (R[0].pivotmin > 0.0 ? -R[0].pivotmin : -999e9);
            for (int me = 1; me < nproc; ++me) {
                c1 = std::max(c1, R[me].amax); c2 = std::max(c2, R[me].pivotmax);
                c3 = std::max(c3, (R[me].pivotmin > 0.0 ? -R[me].pivotmin : -999e9));
            }
            for (auto& K : R) { K.amax = c1; K.pivotmax = c2; K.pivotmin = -c3; if (K.pivotmin == 999e9) K.pivotmin = -1.0; }

            // ---- share blocks to the LEFT (dmrgg.f90:872-958)
            {
                std::vector<std::vector<double>> msg(nproc);
                for (int me = 1; me < nproc; ++me) {
                    Rank& K = R[me]; int q = own[me];
                    if (K.upd[q]) {
                        size_t cnt = (size_t)K.rr[q - 1] * N(q);
                        const double* src = &K.arg[q].at(1, 1, K.r[q]);
                        msg[me].assign(src, src + cnt);
                    }
                }
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer_nt) if (outer_nt > 1)     // receivers are independent (the messages are pre-exchange copies)
                for (int me = 0; me < nproc - 1; ++me) {
                    Rank& K = R[me]; const std::vector<int>& r = K.r; const std::vector<int>& rr = K.rr;
                    int p = own[me + 1] - 1;
                    bool needrecv = K.upd[p + 1];
                    if (!needrecv) continue;
                    const std::vector<double>& arow1 = msg[me + 1];  // (rr(p), n(p+1))
                    if (arow1.size() != (size_t)rr[p] * N(p + 1)) { std::fprintf(stderr, "oracle: LEFT msg size mismatch\n"); std::abort(); }
                    Core na; na.alloc(r[p], N(p + 1), r[p + 1]);
                    std::copy(K.arg[p + 1].a.begin(), K.arg[p + 1].a.begin() + (size_t)r[p] * N(p + 1) * rr[p + 1], na.a.begin());
                    for (int k = 1; k <= N(p + 1); ++k) for (int j = 1; j <= rr[p]; ++j)
                        na.at(j, k, r[p + 1]) = arow1[(j - 1) + (size_t)rr[p] * (k - 1)];
                    K.arg[p + 1] = na;
                    if (K.upd[p]) {
                        int ii = K.vip[p][r[p] - 1][0], jj = K.vip[p][r[p] - 1][1];
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                        for (int k = 1; k <= N(p + 1); ++k)
                            K.arg[p + 1].at(r[p], k, r[p + 1]) = dmrgg_fun(prob, ii, jj, k, r[p + 1], p, l, m, K.vip);
                        for (int k = 1; k <= N(p + 1); ++k) K.amax = std::max(K.amax, std::fabs(K.arg[p + 1].at(r[p], k, r[p + 1])));
                        K.nevalloc += N(p + 1);
                    }
                    Core nr; nr.alloc(r[p], N(p + 1), r[p + 1]);
                    std::copy(K.row[p + 1].a.begin(), K.row[p + 1].a.begin() + (size_t)r[p] * N(p + 1) * rr[p + 1], nr.a.begin());
                    std::vector<double> brow1(&K.arg[p + 1].at(1, 1, r[p + 1]), &K.arg[p + 1].at(1, 1, r[p + 1]) + (size_t)r[p] * N(p + 1));
                    d2_luar(N(p + 1), r[p], K.inv[p].data(), brow1.data());
                    std::copy(brow1.begin(), brow1.end(), nr.a.begin() + (size_t)r[p] * N(p + 1) * rr[p + 1]);
                    K.row[p + 1] = nr;
                }
            }
            // ---- share blocks to the RIGHT (dmrggmp.f90:572-629; absent from dmrgg.f90)
            {
                std::vector<std::vector<double>> msg(nproc);
                for (int me = 0; me < nproc - 1; ++me) {
                    Rank& K = R[me]; int q = own[me + 1] - 1;
                    if (K.upd[q]) {
                        size_t cnt = (size_t)N(q + 1) * K.rr[q + 1];
                        msg[me].resize(cnt);
                        const double* src = &K.arg[q + 1].at(K.r[q], 1, 1);
                        for (size_t x = 0; x < cnt; ++x) msg[me][x] = src[x * (size_t)K.r[q]];
                    }
                }
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer_nt) if (outer_nt > 1)
                for (int me = 1; me < nproc; ++me) {
                    Rank& K = R[me]; const std::vector<int>& r = K.r; const std::vector<int>& rr = K.rr;
                    int p = own[me];
                    bool needrecv = K.upd[p - 1];
                    if (!needrecv) continue;
                    const std::vector<double>& acol1 = msg[me - 1];  // (n(p), rr(p))
                    if (acol1.size() != (size_t)N(p) * rr[p]) { std::fprintf(stderr, "oracle: RIGHT msg size mismatch\n"); std::abort(); }
                    Core na; na.alloc(r[p - 1], N(p), r[p]);
                    for (int k = 1; k <= r[p]; ++k) for (int j = 1; j <= N(p); ++j) for (int i = 1; i <= rr[p - 1]; ++i)
                        na.at(i, j, k) = K.arg[p].a[(size_t)(i - 1) + (size_t)rr[p - 1] * ((j - 1) + (size_t)N(p) * (k - 1))];
                    for (int k = 1; k <= rr[p]; ++k) for (int j = 1; j <= N(p); ++j)
                        na.at(r[p - 1], j, k) = acol1[(j - 1) + (size_t)N(p) * (k - 1)];
                    K.arg[p] = na;
                    if (K.upd[p]) {
                        int kk = K.vip[p][r[p] - 1][2], qq = K.vip[p][r[p] - 1][3];
#pragma omp parallel for schedule(static) num_threads(inner_nt)
                        for (int j = 1; j <= N(p); ++j)
                            K.arg[p].at(r[p - 1], j, r[p]) = dmrgg_fun(prob, r[p - 1], j, kk, qq, p, l, m, K.vip);
                        for (int j = 1; j <= N(p); ++j) K.amax = std::max(K.amax, std::fabs(K.arg[p].at(r[p - 1], j, r[p])));
                        K.nevalloc += N(p);
                    }
                    Core nc; nc.alloc(r[p - 1], N(p), r[p]);
                    std::vector<double> bcol1((size_t)N(p) * r[p]);
                    for (int k = 1; k <= r[p]; ++k) for (int j = 1; j <= N(p); ++j) {
                        for (int i = 1; i <= rr[p - 1]; ++i)
                            nc.at(i, j, k) = K.col[p].a[(size_t)(i - 1) + (size_t)rr[p - 1] * ((j - 1) + (size_t)N(p) * (k - 1))];
                        bcol1[(j - 1) + (size_t)N(p) * (k - 1)] = K.arg[p].at(r[p - 1], j, k);
                    }
                    d2_lual(N(p), r[p], K.inv[p].data(), bcol1.data());
                    for (int k = 1; k <= r[p]; ++k) for (int j = 1; j <= N(p); ++j)
                        nc.at(r[p - 1], j, k) = bcol1[(j - 1) + (size_t)N(p) * (k - 1)];
                    K.col[p] = nc;
                }
            }
        }  // nproc > 1

        for (auto& K : R) K.pivotmax_prev = K.pivotmax;
        nevalall = 0; for (auto& K : R) nevalall += K.nevalloc;

        // ---- report (dmrgg.f90:969-1008)
        double t2 = timef();
        std::snprintf(line, sizeof line, "%3d%2s rank%5.1f time: %s n_evals: %10lld", it, sdir, report_erank(), fmt_e(t2, 9, 3).c_str(), nevalall);
        std::string s = line;
        if (has_quad) {
            std::vector<std::vector<Core>> ttqq(nproc, std::vector<Core>(m + 1));
            std::vector<std::vector<Core>*> ptr(nproc);
#pragma omp parallel for schedule(dynamic, 1) num_threads(outer_nt) if (outer_nt > 1)     // every rank contracts its own cores
            for (int me = 0; me < nproc; ++me) {
                Rank& K = R[me];
                int first = own[me], last = own[me + 1] - 1; if (me == nproc - 1) last = m;
                for (int p = l; p <= m; ++p) ttqq[me][p].alloc(1, 1, 1);
                // ttqq%r(own(me)-1:own(me+1)) = r(...)
                for (int p = l; p <= m; ++p) {
                    int a0 = (p - 1 >= own[me] - 1 && p - 1 <= own[me + 1]) ? K.r[p - 1] : 1;
                    int a1 = (p >= own[me] - 1 && p <= own[me + 1]) ? K.r[p] : 1;
                    ttqq[me][p].alloc(a0, 1, a1);
                }
                for (int p = first; p <= last; ++p) {
                    if (K.arg[p].r0 != K.r[p - 1] || K.arg[p].r1 != K.r[p]) {
                        std::fprintf(stderr, "oracle: core shape mismatch rank %d core %d: (%d,%d) vs r=(%d,%d)\n", me, p, K.arg[p].r0, K.arg[p].r1, K.r[p - 1], K.r[p]);
                        std::abort();
                    }
                    for (int k = 1; k <= K.r[p]; ++k)
                        dgemv_n(K.r[p - 1], N(p), 1.0, &K.arg[p].at(1, 1, k), K.r[p - 1], A.quad[p].data(), 1, 0.0, &ttqq[me][p].at(1, 1, k));
                }
                ptr[me] = &ttqq[me];
            }
            dtt_lua_all(nproc, own, ptr, R);
            val = dtt_quad_all(nproc, own, m, ptr, nullptr);
            if (A.has_tru) s += " err " + fmt_e(std::fabs(1.0 - val / A.tru), 8, 3) + " val " + fmt_e(val, 20, 14);
            else s += " cnv " + fmt_e(std::fabs(1.0 - val / val_prev), 8, 3) + " val " + fmt_e(val, 20, 14);
            val_prev = val;
        }
        res.text += s + "\n";
        if (A.verbose) std::puts(s.c_str());
        res.vals.push_back(val); res.nevals.push_back(nevalall); res.amaxs.push_back(R[0].amax);
        res.pivotmaxs.push_back(R[0].pivotmax); res.eranks.push_back(report_erank()); res.times.push_back(t2);

        // ---- exit conditions (dmrgg.f90:1010-1019); identical on every rank after the allreduce
        if (A.maxrank > 0) ready = ready || (it + 1 >= A.maxrank);
        if (A.accuracy >= 0) {
            if (R[0].pivotmax <= A.accuracy * R[0].amax) strike += 1; else strike = 0;
            ready = ready || (strike >= 3);
        }
    }
    res.nsweeps = it;

    // ---- finalise (dmrgg.f90:1022-1049)
    {
        std::vector<std::vector<Core>*> ptr(nproc);
        for (int me = 0; me < nproc; ++me) ptr[me] = &R[me].arg;
        dtt_lua_all(nproc, own, ptr, R);
        res.cores.resize(m + 1);
        res.ranks.assign(m + 1, 1);
        for (int me = 0; me < nproc; ++me) {
            int first = own[me], last = own[me + 1] - 1; if (me == nproc - 1) last = m;
            for (int p = first; p <= last; ++p) { res.cores[p] = R[me].arg[p]; res.ranks[p - 1] = R[me].arg[p].r0; res.ranks[p] = R[me].arg[p].r1; }
        }
        nevalall = 0; for (auto& K : R) nevalall += K.nevalloc;
        res.neval = nevalall;
        // driver-level dtt_quad(tt, qq) on the finalised train (test_crs_ising.f90:158)
        std::vector<std::vector<Core>*> ptr2(nproc);
        for (int me = 0; me < nproc; ++me) ptr2[me] = &R[me].arg;
        res.quad_final = dtt_quad_all(nproc, own, m, ptr2, has_quad ? &A.quad : nullptr);
    }
    res.seconds = timef();
    return 0;
}

}  // namespace

// =============================================================================
// C interface (ctypes)
// =============================================================================
extern "C" {

void* tto_create(int kind, int d, const int* n, const double* par, long npar, const double* aux, long naux) {
    Oracle* O = new Oracle();
    O->prob.kind = kind; O->prob.d = d;
    O->prob.n.assign(n, n + d);
    O->prob.par.assign(par, par + npar);
    if (aux && naux > 0) O->prob.aux.assign(aux, aux + naux);
    return O;
}
void tto_destroy(void* h) { delete (Oracle*)h; }
void tto_set_rank_concurrency(int on) { g_rank_concurrency = on; }
double tto_visit_seconds(int reset) { double v = g_visit_seconds; if (reset) g_visit_seconds = 0; return v; }
void tto_set_exp_mode(void* h, int mode) { ((Oracle*)h)->prob.exp_mode = mode; }

// quad: concatenated weight vectors sum(n) doubles, or NULL.  own: P+1 ints or NULL.
int tto_run(void* h, int maxrank, double accuracy, int piv, int P, const int* own, const double* quad,
            int has_tru, double tru, unsigned long long seed, int verbose) {
    Oracle* O = (Oracle*)h;
    RunArgs A;
    A.maxrank = maxrank; A.accuracy = accuracy; A.piv = piv; A.P = P;
    if (own) A.own.assign(own, own + P + 1);
    if (quad) {
        A.quad.resize(O->prob.d + 1);
        size_t off = 0;
        for (int p = 1; p <= O->prob.d; ++p) { A.quad[p].assign(quad + off, quad + off + O->prob.n[p - 1]); off += O->prob.n[p - 1]; }
    }
    A.has_tru = has_tru != 0; A.tru = tru; A.seed = seed; A.verbose = verbose;
    return run(*O, A);
}
int tto_nsweeps(void* h) { return ((Oracle*)h)->res.nsweeps; }
long long tto_neval(void* h) { return ((Oracle*)h)->res.neval; }
double tto_seconds(void* h) { return ((Oracle*)h)->res.seconds; }
double tto_quad_final(void* h) { return ((Oracle*)h)->res.quad_final; }
void tto_ranks(void* h, int* out) { auto& r = ((Oracle*)h)->res.ranks; std::copy(r.begin(), r.end(), out); }
// per-sweep series, length nsweeps+1: which = 0 val, 1 neval, 2 amax, 3 pivotmax, 4 erank, 5 time
void tto_series(void* h, int which, double* out) {
    Result& r = ((Oracle*)h)->res;
    for (size_t i = 0; i < r.vals.size(); ++i) {
        switch (which) {
            case 0: out[i] = r.vals[i]; break;
            case 1: out[i] = (double)r.nevals[i]; break;
            case 2: out[i] = r.amaxs[i]; break;
            case 3: out[i] = r.pivotmaxs[i]; break;
            case 4: out[i] = r.eranks[i]; break;
            default: out[i] = r.times[i]; break;
        }
    }
}
long tto_pivlog_count(void* h) { return (long)((Oracle*)h)->res.pivlog.size(); }
// ints: count x 8 row-major {it, vrank, bond, ii, jj, kk, qq, upd}; vals: count pivots
void tto_pivlog(void* h, int* ints, double* vals) {
    auto& L = ((Oracle*)h)->res.pivlog;
    for (size_t i = 0; i < L.size(); ++i) {
        int* o = ints + 8 * i;
        o[0] = L[i].it; o[1] = L[i].vrank; o[2] = L[i].bond; o[3] = L[i].ii; o[4] = L[i].jj; o[5] = L[i].kk; o[6] = L[i].qq; o[7] = L[i].upd;
        vals[i] = L[i].pivot;
    }
}
// finalised core k (1-based), packed (r(k-1), n(k), r(k)) column-major
void tto_core(void* h, int k, double* out) { auto& c = ((Oracle*)h)->res.cores[k]; std::copy(c.a.begin(), c.a.end(), out); }
long tto_text(void* h, char* buf, long cap) {
    auto& t = ((Oracle*)h)->res.text;
    if (buf && cap > 0) { long c = std::min<long>(cap - 1, (long)t.size()); std::memcpy(buf, t.data(), c); buf[c] = 0; }
    return (long)t.size();
}

// ---- unit-level exports
double tto_integrand(void* h, const int* ind) { return fun(((Oracle*)h)->prob, ind); }
void tto_lgwt(int n, double* x, double* w) { lgwt(n, x, w); }
void tto_share(int first, int last, int nproc, int* own) { share(first, last, nproc, own); }
void tto_lottery2(int npnt, int m, int n, const double* wcol, const double* wrow, const double* d, int* points) { lottery2(npnt, m, n, wcol, wrow, d, points); }
double tto_stream_uniform(unsigned long long seed, int vrank, unsigned long long k) { return stream_uniform(seed, vrank, k); }
int tto_idamax(long n, const double* x) { return idamax(n, x); }
void tto_d2_lual(long m, int r, const double* g, double* col, int from) { d2_lual(m, r, g, col, from); }
void tto_d2_luar(long n, int r, const double* g, double* row, int from) { d2_luar(n, r, g, row, from); }
double tto_erank(int d, const int* n, const int* r) {
    std::vector<int> nv(n, n + d), rv(r, r + d + 1);
    return erank(1, d, nv, rv);
}
// ort0_d (lib/ort.f90:17-81) = LAPACK dgeqrf + dorgqr.  LAPACK is a third-party dependency that the reference links
// unpinned ("-llapack", Makefile:18) and that is absent from /root/reference; this restates its published unblocked
// algorithm (netlib dgeqr2 / dlarfg / dlarf / dorg2r: beta = -sign(alpha) * norm, v(1) = 1, tau = (beta - alpha) / beta).
// PINNED in tests/test_qr.py against numpy.linalg.qr, which calls the same LAPACK routines (dgeqrf + dorgqr).
// a: m x n column-major (lda = m), overwritten?  no: inputs are copied.  q: m x n, r: n x n (zeros below the diagonal).
void tto_qr_thin(int m, int n, const double* a, double* q, double* r) {
    if (m < n) {     // ort.f90:32-46
        for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) r[i + (size_t)n * j] = (i < m) ? a[i + (size_t)m * j] : 0.0;
        for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) q[i + (size_t)m * j] = (i == j) ? 1.0 : 0.0;
        return;
    }
    std::vector<double> y(a, a + (size_t)m * n), tau(n, 0.0);
    auto Y = [&](int i, int j) -> double& { return y[i + (size_t)m * j]; };
    for (int k = 0; k < n; ++k) {                       // dgeqr2
        double ssq = 0.0;
        for (int i = k + 1; i < m; ++i) ssq += Y(i, k) * Y(i, k);
        const double alpha = Y(k, k);
        double beta = alpha; tau[k] = 0.0;
        if (ssq != 0.0) {                               // dlarfg
            beta = -std::copysign(std::sqrt(alpha * alpha + ssq), alpha);
            tau[k] = (beta - alpha) / beta;
            const double sc = 1.0 / (alpha - beta);
            for (int i = k + 1; i < m; ++i) Y(i, k) *= sc;
        }
        for (int j = k + 1; j < n; ++j) {               // dlarf from the left, v(k) = 1
            double w = Y(k, j);
            for (int i = k + 1; i < m; ++i) w += Y(i, k) * Y(i, j);
            Y(k, j) -= tau[k] * w;
            for (int i = k + 1; i < m; ++i) Y(i, j) -= tau[k] * w * Y(i, k);
        }
        Y(k, k) = beta;
    }
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) r[i + (size_t)n * j] = (i <= j) ? Y(i, j) : 0.0;
    for (int k = n - 1; k >= 0; --k) {                  // dorg2r
        if (k < n - 1) {
            for (int j = k + 1; j < n; ++j) {
                double w = Y(k, j);
                for (int i = k + 1; i < m; ++i) w += Y(i, k) * Y(i, j);
                Y(k, j) -= tau[k] * w;
                for (int i = k + 1; i < m; ++i) Y(i, j) -= tau[k] * w * Y(i, k);
            }
        }
        for (int i = k + 1; i < m; ++i) Y(i, k) = -tau[k] * Y(i, k);
        Y(k, k) = 1.0 - tau[k];
        for (int i = 0; i < k; ++i) Y(i, k) = 0.0;
    }
    std::copy(y.begin(), y.end(), q);
}
// dtt_ort (lib/tt.f90:130-198): orthogonalise a train from the left.  cores: the d cores concatenated, core k is
// r(k-1) x n(k) x r(k) column-major; overwritten.  Only the rank-preserving case r(k) <= r(k-1) n(k) (every cross result).
// dgeqrf/dorgqr = tto_qr_thin above; dnrm2 as a plain root of the sum of squares; dgemm 'n','n' in reference order.
int tto_tt_ort(int d, const int* n, const int* r, double* cores) {
    std::vector<size_t> off(d + 1, 0);
    for (int k = 0; k < d; ++k) off[k + 1] = off[k] + (size_t)r[k] * n[k] * r[k + 1];
    double lognrm = 0.0;
    for (int k = 0; k + 1 < d; ++k) {
        const int mm = r[k] * n[k], nn = r[k + 1];
        const long kk = (long)n[k + 1] * r[k + 2];
        if (mm < nn) return 1;
        std::vector<double> q((size_t)mm * nn), mat((size_t)nn * nn);
        tto_qr_thin(mm, nn, cores + off[k], q.data(), mat.data());
        double ss = 0.0;
        for (double x : mat) ss += x * x;
        const double nrm = std::sqrt(ss);
        if (nrm != 0.0) { const double sc = 1.0 / nrm; for (double& x : mat) x = sc * x; lognrm = lognrm + std::log(nrm); }
        std::copy(q.begin(), q.end(), cores + off[k]);
        std::vector<double> u((size_t)nn * kk, 0.0);
        const double* b = cores + off[k + 1];
        for (long c = 0; c < kk; ++c)
            for (int l = 0; l < nn; ++l) { const double t = b[l + (size_t)nn * c]; for (int i = 0; i < nn; ++i) u[i + (size_t)nn * c] += t * mat[i + (size_t)nn * l]; }
        std::copy(u.begin(), u.end(), cores + off[k + 1]);
    }
    {
        double* last = cores + off[d - 1];
        const size_t cnt = off[d] - off[d - 1];
        double ss = 0.0;
        for (size_t x = 0; x < cnt; ++x) ss += last[x] * last[x];
        const double nrm = std::sqrt(ss);
        if (nrm != 0.0) { const double sc = 1.0 / nrm; for (size_t x = 0; x < cnt; ++x) last[x] = sc * last[x]; lognrm = lognrm + std::log(nrm); }
    }
    const double nrm = std::exp(lognrm / d);
    for (size_t x = 0; x < off[d]; ++x) cores[x] = nrm * cores[x];
    return 0;
}
// ztt_quad (lib/dmrgg.f90:1418-1523), one rank: cores real (the drivers give the complex train zero imaginary parts),
// weights complex.  zgemv 'n' per slice (:1469-1471), zgemm 'n','n' chain (:1481-1483), reference BLAS orders.
void tto_quad_complex(int d, const int* n, const int* r, const double* cores, const double* wre, const double* wim, double* out) {
    std::vector<double> pre(1, 1.0), pim(1, 0.0);
    size_t off = 0, woff = 0;
    for (int p = 0; p < d; ++p) {
        const int r0 = r[p], r1 = r[p + 1], np_ = n[p];
        std::vector<double> nre(r1, 0.0), nim(r1, 0.0);
        for (int k = 0; k < r1; ++k) {
            std::vector<double> cre(r0, 0.0), cim(r0, 0.0);
            for (int j = 0; j < np_; ++j)
                for (int i = 0; i < r0; ++i) { const double v = cores[off + i + (size_t)r0 * (j + (size_t)np_ * k)]; cre[i] = cre[i] + wre[woff + j] * v; cim[i] = cim[i] + wim[woff + j] * v; }
            double are = 0.0, aim = 0.0;
            for (int l = 0; l < r0; ++l) { are = are + (cre[l] * pre[l] - cim[l] * pim[l]); aim = aim + (cre[l] * pim[l] + cim[l] * pre[l]); }
            nre[k] = are; nim[k] = aim;
        }
        pre.swap(nre); pim.swap(nim);
        off += (size_t)r0 * np_ * r1; woff += np_;
    }
    out[0] = pre[0]; out[1] = pim[0];
}
// dtt_svd (lib/tt.f90:307-368) with d_svd / chop (lib/mat.f90:340-385, 433-455).  LAPACK dgesvd is an unpinned third-party
// dependency of the reference; its published contract (M = U diag(s) V^T, s descending) is restated with a one-sided Jacobi
// SVD.  PINNED in tests/test_tt_svd.py against numpy.linalg.svd (LAPACK) through ranks and the rounded tensor.
// r: ranks in/out (d+1); cores in/out (compacted to the new ranks); tol < 0 / rmax <= 0: absent.
static void jacobi_svd_rows(int mm, long nn, std::vector<double>& M /*mm x nn col-major*/, std::vector<double>& U, std::vector<double>& sv, std::vector<double>& V) {
    // work on A = M^T (nn x mm): A J = B (orthogonal columns) => M = J diag-normalised...: M^T = B J^T, so M = J B^T = (J) (S Vr) with U = J
    std::vector<double> A((size_t)nn * mm), J((size_t)mm * mm, 0.0);
    for (long c = 0; c < nn; ++c) for (int i = 0; i < mm; ++i) A[c + (size_t)nn * i] = M[i + (size_t)mm * c];
    for (int i = 0; i < mm; ++i) J[i + (size_t)mm * i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rot = false;
        for (int p = 0; p < mm - 1; ++p)
            for (int q = p + 1; q < mm; ++q) {
                double a = 0, b = 0, g = 0;
                for (long c = 0; c < nn; ++c) { const double x = A[c + (size_t)nn * p], y = A[c + (size_t)nn * q]; a += x * x; b += y * y; g += x * y; }
                if (g == 0.0 || std::fabs(g) <= 1e-15 * std::sqrt(a * b)) continue;
                rot = true;
                const double zeta = (b - a) / (2.0 * g), t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (long c = 0; c < nn; ++c) { const double x = A[c + (size_t)nn * p], y = A[c + (size_t)nn * q]; A[c + (size_t)nn * p] = cs * x - sn * y; A[c + (size_t)nn * q] = sn * x + cs * y; }
                for (int i = 0; i < mm; ++i) { const double x = J[i + (size_t)mm * p], y = J[i + (size_t)mm * q]; J[i + (size_t)mm * p] = cs * x - sn * y; J[i + (size_t)mm * q] = sn * x + cs * y; }
            }
        if (!rot) break;
    }
    std::vector<double> nrm(mm); std::vector<int> ord(mm);
    for (int j = 0; j < mm; ++j) { double a = 0; for (long c = 0; c < nn; ++c) a += A[c + (size_t)nn * j] * A[c + (size_t)nn * j]; nrm[j] = std::sqrt(a); ord[j] = j; }
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return nrm[x] > nrm[y]; });
    U.assign((size_t)mm * mm, 0.0); V.assign((size_t)mm * nn, 0.0); sv.assign(mm, 0.0);
    for (int jo = 0; jo < mm; ++jo) {
        const int j = ord[jo]; sv[jo] = nrm[j];
        for (int i = 0; i < mm; ++i) U[i + (size_t)mm * jo] = J[i + (size_t)mm * j];
        for (long c = 0; c < nn; ++c) V[jo + (size_t)mm * c] = nrm[j] > 0 ? A[c + (size_t)nn * j] / nrm[j] : 0.0;     // V(jo, c): rows of V^T... stored mm x nn
    }
}
int tto_tt_ort(int d, const int* n, const int* r, double* cores);
int tto_tt_svd(int d, const int* n, int* r, double* cores, double tol, int rmax) {
    if (tto_tt_ort(d, n, r, cores)) return 1;
    std::vector<std::vector<double>> c(d);
    { size_t off = 0; for (int k = 0; k < d; ++k) { const size_t sz = (size_t)r[k] * n[k] * r[k + 1]; c[k].assign(cores + off, cores + off + sz); off += sz; } }
    double lognrm = 0.0;
    for (int k = d - 1; k >= 1; --k) {
        const int mm = r[k]; const long nn = (long)n[k] * r[k + 1], kk = (long)r[k - 1] * n[k - 1];
        std::vector<double> U, sv, V;
        jacobi_svd_rows(mm, nn, c[k], U, sv, V);
        // chop (mat.f90:433-455)
        int rr = mm; double er2 = 0.0;
        if (rmax > 0 && rmax < rr) { for (int i = rmax; i < rr; ++i) er2 += sv[i] * sv[i]; rr = rmax; }
        if (tol >= 0) {
            double ss = 0; for (double x : sv) ss += x * x;
            const double nrm = std::sqrt(ss), bound = tol * tol * nrm * nrm;
            double er = er2 + sv[rr - 1] * sv[rr - 1];
            while (er < bound && rr > 1) { er2 = er; rr = rr - 1; er = er + sv[rr - 1] * sv[rr - 1]; }
        }
        double ss = 0; for (int j = 0; j < rr; ++j) ss += sv[j] * sv[j];
        const double nrm = std::sqrt(ss);
        std::vector<double> sn(sv.begin(), sv.begin() + rr);
        if (nrm != 0.0) { for (double& x : sn) x = (1.0 / nrm) * x; lognrm = lognrm + std::log(nrm); }
        std::vector<double> left((size_t)kk * rr, 0.0), right((size_t)rr * nn);
        for (int j = 0; j < rr; ++j)
            for (int l = 0; l < mm; ++l) { const double t = U[l + (size_t)mm * j] * sn[j]; for (long x = 0; x < kk; ++x) left[x + (size_t)kk * j] += t * c[k - 1][x + (size_t)kk * l]; }
        for (long cc = 0; cc < nn; ++cc) for (int i = 0; i < rr; ++i) right[i + (size_t)rr * cc] = V[i + (size_t)mm * cc];
        c[k - 1].swap(left); c[k].swap(right); r[k] = rr;
    }
    { double ss = 0; for (double x : c[0]) ss += x * x; const double nrm = std::sqrt(ss); if (nrm != 0.0) { for (double& x : c[0]) x = (1.0 / nrm) * x; lognrm = lognrm + std::log(nrm); } }
    const double sc = std::exp(lognrm / d);
    size_t off = 0;
    for (int k = 0; k < d; ++k) { for (double x : c[k]) cores[off++] = sc * x; }
    return 0;
}
int tto_fmt_e(double v, int w, int dgt, char* buf) { std::string s = fmt_e(v, w, dgt); std::memcpy(buf, s.c_str(), s.size() + 1); return (int)s.size(); }
int tto_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void tto_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

}  // extern "C"
