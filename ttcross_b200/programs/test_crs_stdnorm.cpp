// C++ twin of test_crs_stdnorm.f90 (CLI: DIM N RANK PIV); setup follows test_crs_stdnorm.f90:23-131.
#include "driver_common.hpp"

int main(int argc, char** argv) {
    int d = drv::arg_i(argc, argv, 1, 6), n = drv::arg_i(argc, argv, 2, 65), r = drv::arg_i(argc, argv, 3, 20), piv = drv::arg_i(argc, argv, 4, 1);
    int adj = 0;
    if (n % 2 == 0) { n += 1; adj = 1; }
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    std::printf(" Hi, this is TT cross interpolation for computing integrals...\n");
    std::printf("   dimension:%10d\n", d);
    drv::banner_common(n, adj, r, piv, nparts);
    const double acc = 5 * 2.220446049250313e-16, a = -10.0, b = 10.0;
    const double tru = std::pow(std::sqrt(3.141592653589793238), d);
    std::vector<double> x(n), w(n), par(2 * n);
    ttc_lgwt(n, x.data(), w.data());
    for (int i = 0; i < n; ++i) { par[i] = 0.5 * ((b - a) * x[i] + (a + b)); par[n + i] = (0.5 * (b - a)) * w[i]; }
    std::vector<int> nn(d, n);
    std::vector<double> quad;
    for (int p = 0; p < d; ++p) quad.insert(quad.end(), par.begin() + n, par.end());
    ttc_handle* h = nullptr;
    int st = ttc_create(&h, TTC_STDNORM, d, nn.data(), par.data(), (long)par.size(), nullptr, 0);
    if (st) drv::die(nullptr, st, "ttc_create");
    ttc_set_quad(h, quad.data());
    ttc_set_tru(h, 1, tru);
    int rc = drv::run_and_report(h, r, acc, piv, tru, true, -1);
    ttc_destroy(h);
    return rc;
}
