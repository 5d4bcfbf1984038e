// C++ twin of test_crs_chf.f90 (CLI: DIM N RANK PIV): cross of the MVN density WITHOUT quad/tru (test_crs_chf.f90:122-123),
// then the characteristic function of mean_j exp(X_j) at omega_k = k*pi/300, k = 0..31, as 32 complex quadratures of
// the same train (test_crs_chf.f90:153-168) -- here one batched ttc_quad_complex launch.  With DIM = 4 the values reproduce
// the reference's own table get_reference_val (test_crs_chf.f90:232-271); pass its values through TTC_CHF_TABLE
// (tests/golden/reference/chf_table.json flattened to "re im" lines) to print the 'analytic value' / 'correct digits' lines.
#include "driver_common.hpp"
#include <complex>
#include <fstream>

int main(int argc, char** argv) {
    int d = drv::arg_i(argc, argv, 1, 6), n = drv::arg_i(argc, argv, 2, 65), r = drv::arg_i(argc, argv, 3, 20), piv = drv::arg_i(argc, argv, 4, 1);
    int adj = 0;
    if (n % 2 == 0) { n += 1; adj = 1; }
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    std::printf(" Hi, this is TT cross interpolation for computing integrals...\n");
    std::printf("   dimension:%10d\n", d);
    std::printf("   quadratur:%10d", n);                                   // the reference prints this field twice (:60-61)
    drv::banner_common(n, adj, r, piv, nparts);
    const double acc = 500 * 2.220446049250313e-16;
    const double a = (double)0.525170f, b = (double)8.525170f;             // single-precision literals (:86-87)
    const double pi = 3.141592653589793;
    std::printf("   Computing quadrature weights...\n");
    std::vector<double> x(n), w(n), par(2 * n);
    ttc_lgwt(n, x.data(), w.data());
    for (int i = 0; i < n; ++i) { par[i] = 0.5 * ((b - a) * x[i] + (a + b)); par[n + i] = (0.5 * (b - a)) * w[i]; }
    std::vector<double> aux = drv::mvn_aux(d, 0.0, 1.0);
    std::vector<int> nn(d, n);
    ttc_handle* h = nullptr;
    int st = ttc_create(&h, TTC_MVN, d, nn.data(), par.data(), (long)par.size(), aux.data(), (long)aux.size());
    if (st) drv::die(nullptr, st, "ttc_create");
    if (nparts > 1) { st = ttc_set_partition(h, nparts, nullptr); if (st) drv::die(h, st, "ttc_set_partition"); }
    if (std::getenv("TTC_SEED")) ttc_set_seed(h, std::strtoull(std::getenv("TTC_SEED"), nullptr, 10));
    ttc_set_verbose(h, std::getenv("TTC_QUIET") ? 0 : 1);
    std::printf("   Running TT-cross...\n");
    auto t1 = std::chrono::steady_clock::now();
    st = ttc_dmrgg(h, r, acc, piv);
    if (st) { std::printf("%s\n", ttc_last_error(h)); return 1; }
    double tcrs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    std::printf("...with%12lld evaluations completed in %s sec.\n", ttc_neval(h), drv::fmt_e(tcrs, 12, 4).c_str());
    std::printf("   Preparing quadrature tensor...\n");
    const int K = 32;
    size_t tot = (size_t)d * n;
    std::vector<double> wre(K * tot), wim(K * tot), ore(K), oim(K);
    for (int k = 0; k < K; ++k) {
        double omega = k * pi / (300.0 - 0.0);
        for (int p = 0; p < n; ++p) {
            std::complex<double> wc = std::exp(std::complex<double>(0.0, 1.0) * omega * std::exp(par[p]) / (double)d);
            std::complex<double> q = std::complex<double>(par[n + p], 0.0) * wc;
            for (int i = 0; i < d; ++i) { wre[k * tot + (size_t)i * n + p] = q.real(); wim[k * tot + (size_t)i * n + p] = q.imag(); }
        }
    }
    st = ttc_quad_complex(h, K, wre.data(), wim.data(), ore.data(), oim.data());
    if (st) drv::die(h, st, "ttc_quad_complex");
    std::vector<std::complex<double>> table;
    if (const char* tp = std::getenv("TTC_CHF_TABLE")) {
        std::ifstream f(tp);
        double re, im;
        while (f >> re >> im) table.emplace_back((double)(float)re, (double)(float)im);    // default-kind cmplx() (:238-269)
    }
    for (int k = 0; k < K; ++k) {
        std::printf("computed value: %s%s\n", drv::fmt_e(ore[k], 50, 40).c_str(), drv::fmt_e(oim[k], 50, 40).c_str());
        if ((int)table.size() == K) {
            std::complex<double> tru = table[k], ans(ore[k], oim[k]);
            std::printf("analytic value: %s%s\n", drv::fmt_e(tru.real(), 50, 40).c_str(), drv::fmt_e(tru.imag(), 50, 40).c_str());
            std::printf("correct digits:%7.2f\n", -std::log(std::abs(std::complex<double>(1.0, 0.0) - ans / tru)) / std::log(10.0));
        }
    }
    std::printf("Good bye.\n");
    ttc_destroy(h);
    return 0;
}
