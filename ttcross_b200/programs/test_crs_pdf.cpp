// C++ twin of test_crs_pdf.f90 (CLI: DIM N RANK PIV): the characteristic-function pipeline of test_crs_chf.f90 (MVN cross
// without quad/tru, :128-129; 32 complex quadratures, :153-168 -- one batched ttc_quad_complex launch here), then the COS
// series of the density of mean_j exp(X_j) on 200 points of [0, 300] (cos_approximate_array, lib/cos_approx.f90:90-127)
// written as two es25.17 columns (:186-199).  Output file: ./out/tt-cross-pdf.txt like the reference, or $TTC_PDF_OUT.
// The reference then shells out to its matplotlib scripts (:206-214); this twin stops at the file.  With $TTC_TT_OUT set it
// also stores the train like test_crs_store.f90:136 does (in the reference's TT stream format instead of HDF5); the binary
// test_crs_store is this file compiled with -DTTC_STORE_TWIN, which makes that step unconditional (./out/tensor_train.tt).
#include "driver_common.hpp"
#include <complex>

int main(int argc, char** argv) {
    int d = drv::arg_i(argc, argv, 1, 6), n = drv::arg_i(argc, argv, 2, 65), r = drv::arg_i(argc, argv, 3, 20), piv = drv::arg_i(argc, argv, 4, 1);
    int adj = 0;
    if (n % 2 == 0) { n += 1; adj = 1; }
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    std::printf(" Hi, this is TT cross interpolation for computing integrals...\n");
    std::printf("   dimension:%10d\n", d);
    drv::banner_common(n, adj, r, piv, nparts);
    const double acc = 500 * 2.220446049250313e-16;
    const double a = (double)0.525170f, b = (double)8.525170f;             // single-precision literals (:95-96)
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
    std::printf("   Calculating phis...\n");
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
    // test_crs_store.f90:136 saves the train right after the cross ("out/tensor_train.h5", HDF5 there; the TT stream format of
    // lib/ttio.f90 here -- no HDF5 library in this build).  Built with -DTTC_STORE_TWIN this file IS the twin of
    // test_crs_store.f90 (the same pipeline as test_crs_pdf.f90 plus the store step): the train is always written.
    const char* tt_out = std::getenv("TTC_TT_OUT");
#ifdef TTC_STORE_TWIN
    if (!tt_out) tt_out = "./out/tensor_train.tt";
#endif
    if (tt_out) {
        st = ttc_write(h, tt_out);
        if (st) drv::die(h, st, "ttc_write");
        std::printf("   Train written to: %s\n", tt_out);
    }
    ttc_destroy(h);
    std::printf("   Phi values computed.\n");
    std::vector<std::complex<double>> phis(K);
    for (int k = 0; k < K; ++k) phis[k] = std::complex<double>(ore[k], oim[k]);
    const int n_pts = 200;
    std::vector<double> xs(n_pts), pdf;
    for (int i = 1; i <= n_pts; ++i) xs[i - 1] = 0.0 + (300.0 - 0.0) * (i - 1) / (n_pts - 1);
    drv::cos_approximate_array(xs, phis, 0.0, 300.0, 32, pdf);
    const char* fn = std::getenv("TTC_PDF_OUT") ? std::getenv("TTC_PDF_OUT") : "./out/tt-cross-pdf.txt";
    std::FILE* f = std::fopen(fn, "w");
    if (!f) std::printf(" Error opening file: %s\n", fn);
    else {
        std::printf(" Writing PDF output to: %s\n", fn);
        for (int i = 0; i < n_pts; ++i) std::fprintf(f, "%25.17E %25.17E\n", xs[i], pdf[i]);
        std::fclose(f);
    }
    return 0;
}
