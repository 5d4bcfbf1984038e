// C++ twin of test_crs_mvn.f90 (CLI: DIM N RANK PIV); setup follows test_crs_mvn.f90:23-133 and lib/mvn_pdf.f90:15-111.
#include "driver_common.hpp"

int main(int argc, char** argv) {
    int d = drv::arg_i(argc, argv, 1, 6), n = drv::arg_i(argc, argv, 2, 65), r = drv::arg_i(argc, argv, 3, 20), piv = drv::arg_i(argc, argv, 4, 1);
    int adj = 0;
    if (n % 2 == 0) { n += 1; adj = 1; }
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    std::printf(" Hi, this is TT cross interpolation for computing integrals...\n");
    std::printf("   dimension:%10d\n", d);
    drv::banner_common(n, adj, r, piv, nparts);
    const double acc = 500 * 2.220446049250313e-16;
    const double a = (double)0.525170f, b = (double)8.525170f;     // single-precision literals in the reference
    const double tru = 1.0;
    std::vector<double> x(n), w(n), par(2 * n);
    ttc_lgwt(n, x.data(), w.data());
    for (int i = 0; i < n; ++i) { par[i] = 0.5 * ((b - a) * x[i] + (a + b)); par[n + i] = (0.5 * (b - a)) * w[i]; }
    std::vector<double> aux = drv::mvn_aux(d, 0.0, 1.0);            // mvn_init(d, 0, 1)
    std::vector<int> nn(d, n);
    std::vector<double> quad;
    for (int p = 0; p < d; ++p) quad.insert(quad.end(), par.begin() + n, par.end());
    ttc_handle* h = nullptr;
    int st = ttc_create(&h, TTC_MVN, d, nn.data(), par.data(), (long)par.size(), aux.data(), (long)aux.size());
    if (st) drv::die(nullptr, st, "ttc_create");
    ttc_set_quad(h, quad.data());
    ttc_set_tru(h, 1, tru);
    int rc = drv::run_and_report(h, r, acc, piv, tru, true, -1);
    ttc_destroy(h);
    return rc;
}
