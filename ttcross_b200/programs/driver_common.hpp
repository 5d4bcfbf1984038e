// Shared pieces of the C++ twins of the reference driver programs (test_crs_ising.f90, test_crs_mvn.f90,
// test_crs_stdnorm.f90): same positional CLI (readarg, lib/default.f90:40-78), same banner, same final lines.
// The Fortran originals cannot be compiled in this image (no Fortran compiler); fortran/ holds the ISO_C_BINDING
// versions for sites that have one.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <chrono>
#include <complex>

#include "../../include/ttcross_b200.h"

namespace drv {

inline int arg_i(int argc, char** argv, int pos, int def) { return (pos < argc && argv[pos][0]) ? std::atoi(argv[pos]) : def; }
inline char arg_a(int argc, char** argv, int pos, char def) { return (pos < argc && argv[pos][0]) ? argv[pos][0] : def; }

inline std::string fmt_e(double v, int w, int dgt) {
    char buf[128];
    std::string s;
    if (v == 0.0) s = "0." + std::string(dgt, '0') + "E+00";
    else {
        std::snprintf(buf, sizeof buf, "%.*e", dgt - 1, std::fabs(v));
        std::string t = buf;
        size_t epos = t.find('e');
        int ex = std::atoi(t.c_str() + epos + 1) + 1;
        std::string digits;
        for (size_t i = 0; i < epos; ++i) if (t[i] != '.') digits += t[i];
        s = std::string(v < 0 ? "-" : "") + "0." + digits;
        char eb[16];
        if (std::abs(ex) < 100) std::snprintf(eb, sizeof eb, "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
        else std::snprintf(eb, sizeof eb, "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
        s += eb;
        if ((int)s.size() > w) { size_t z = (s[0] == '-') ? 1 : 0; if (s[z] == '0') s.erase(z, 1); }
    }
    if ((int)s.size() > w) s = std::string(w, '*');
    if ((int)s.size() < w) s = std::string(w - s.size(), ' ') + s;
    return s;
}

inline void banner_common(int n, int adj, int r, int piv, int nparts) {
    if (adj == 0) std::printf("   quadratur:%10d\n", n); else std::printf("   quadratur:%10d (adjusted)\n", n);
    std::printf("   TT ranks :%10d\n", r);
    std::printf("   pivoting :%10d\n", piv);
    std::printf("   MPI procs:%10d\n", nparts);          // virtual partitions take the place of MPI ranks
    std::printf("   OMP thrds:%10d\n", 1);
    std::printf("   sizeof(d):%10d\n", 64);
    std::printf("   epsilon  :%s\n", fmt_e(2.220446049250313e-16, 10, 3).c_str());
}

inline void die(ttc_handle* h, int st, const char* what) {
    std::fprintf(stderr, "%s failed (%d): %s\n", what, st, ttc_last_error(h));
    std::exit(1);
}

// Runs the cross and prints the reference's closing lines.  rescale_pow >= 0 prints ' / (5**k)' like test_crs_ising.f90:160-161.
inline int run_and_report(ttc_handle* h, int maxrank, double acc, int piv, double tru, bool has_tru, int rescale_pow) {
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    if (nparts > 1) { int st = ttc_set_partition(h, nparts, nullptr); if (st) die(h, st, "ttc_set_partition"); }
    if (std::getenv("TTC_SEED")) ttc_set_seed(h, std::strtoull(std::getenv("TTC_SEED"), nullptr, 10));
    ttc_set_verbose(h, std::getenv("TTC_QUIET") ? 0 : 1);
    auto t1 = std::chrono::steady_clock::now();
    int st = ttc_dmrgg(h, maxrank, acc, piv);
    if (st) { std::printf("%s\n", ttc_last_error(h)); return 1; }   // write(*,*) msg; stop
    double tcrs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    if (std::getenv("TTC_QUIET")) { std::vector<char> buf(ttc_text(h, nullptr, 0) + 1); ttc_text(h, buf.data(), (long)buf.size()); std::fputs(buf.data(), stdout); }
    std::printf("...with%12lld evaluations completed in %s sec.\n", ttc_neval(h), fmt_e(tcrs, 12, 4).c_str());
    double val = 0;
    st = ttc_quad(h, &val);
    if (st) die(h, st, "ttc_quad");
    if (rescale_pow >= 0) std::printf("computed value:%s / (5**%4d)\n", fmt_e(val, 50, 40).c_str(), rescale_pow);
    else std::printf("computed value:%s\n", fmt_e(val, 50, 40).c_str());
    if (has_tru) {
        std::printf("analytic value:%s\n", fmt_e(tru, 50, 40).c_str());
        std::printf("correct digits:%7.2f\n", -std::log(std::fabs(1.0 - val / tru)) / std::log(10.0));
    }
    std::printf("Good bye.\n");
    return 0;
}

// Gauss-Jordan with partial pivoting (the reference calls LAPACK dgetrf/dgetri; any correct inverse is input data)
inline void inv_det(std::vector<double> a, int n, std::vector<double>& inv, double& det) {
    std::vector<double> m((size_t)n * 2 * n, 0.0);
    auto M = [&](int i, int j) -> double& { return m[(size_t)i * 2 * n + j]; };
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) M(i, j) = a[(size_t)i * n + j]; M(i, n + i) = 1.0; }
    det = 1.0;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(M(i, k)) > std::fabs(M(piv, k))) piv = i;
        if (piv != k) { for (int j = 0; j < 2 * n; ++j) std::swap(M(k, j), M(piv, j)); det = -det; }
        det = det * M(k, k);
        double pk = M(k, k);
        for (int j = 0; j < 2 * n; ++j) M(k, j) = M(k, j) / pk;
        for (int i = 0; i < n; ++i) {
            if (i == k || M(i, k) == 0.0) continue;
            double f = M(i, k);
            for (int j = 0; j < 2 * n; ++j) M(i, j) = M(i, j) - f * M(k, j);
        }
    }
    inv.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) inv[i + (size_t)j * n] = M(i, n + j);   // column-major
}
inline double powi(double x, int m) {   // libgcc __powidf2 (what gfortran emits for real**integer)
    unsigned n = m < 0 ? -m : m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return m < 0 ? 1.0 / y : y;
}

// aux blob of the MVN integrand = module state of lib/mvn_pdf.f90:4-11 after mvn_init(n, r, T) (mvn_pdf.f90:15-60)
inline std::vector<double> mvn_aux(int d, double rr, double T) {
    const double sigma = 0.4, corr = 0.5;
    std::vector<double> cov((size_t)d * d), inv;
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) cov[(size_t)i * d + j] = ((i == j) ? sigma * sigma : sigma * corr * sigma) * T;
    double det;
    inv_det(cov, d, inv, det);
    std::vector<double> aux(d, std::log(100.0) + (rr - 0.5 * (sigma * sigma)) * T);
    aux.insert(aux.end(), inv.begin(), inv.end());
    aux.push_back(std::sqrt(powi(2.0 * 3.141592653589793, d) * det));
    return aux;
}

// cos_approximate_array (lib/cos_approx.f90:90-127): pdf(x) = sum_k' coeff_k cos(omega_k (x - a)), omega_k = k pi / (b - a),
// coeff_k = 2 / (b - a) Re(phi_k exp(-i omega_k a)), the k = 0 term halved; n_terms > size(phis) prints the reference's
// error line and returns zeros (:107-111)
inline void cos_approximate_array(const std::vector<double>& xs, const std::vector<std::complex<double>>& phis, double lower_bound,
                                  double upper_bound, int n_terms, std::vector<double>& pdf_vals) {
    pdf_vals.assign(xs.size(), 0.0);
    if (n_terms > (int)phis.size()) { std::printf(" Error: n_terms exceeds the size of phis.\n"); return; }
    const double pi = 3.1415926535897932384626433832795;
    const double pi_over_bound = pi / (upper_bound - lower_bound);
    for (int k = 0; k < n_terms; ++k) {
        const double omega = k * pi_over_bound;
        double coeff = 2.0 / (upper_bound - lower_bound) * (phis[k] * std::exp(-1.0 * std::complex<double>(0.0, 1.0) * omega * lower_bound)).real();
        if (k == 0) coeff = coeff / 2.0;
        for (size_t i = 0; i < xs.size(); ++i) pdf_vals[i] = pdf_vals[i] + coeff * std::cos(omega * (xs[i] - lower_bound));
    }
}

}  // namespace drv
