// C++ twin of the reference program test_crs_ising.f90 (same CLI: KIND INDEX N RANK PIV), driving the CUDA sweep
// through the C-ABI.  Setup follows test_crs_ising.f90:25-153.
#include "driver_common.hpp"
#include <map>

static double ising_tru(char a, int m) {   // test_crs_ising.f90:71-100, rounded to double
    const double tpi = 6.2831853071795864769, log2 = 0.69314718055994530942, zeta3 = 1.2020569031595942854, c3 = 0.78130241289648629687;
    if (a == 'c') {
        static const std::map<int, double> t = {{2, 1.0}, {3, c3}, {4, 0.70119986017642999982}, {5, 0.66575980019993742832},
            {6, 0.64863420903100707526}, {8, 0.63548402675916322614}, {16, 0.63050394617323726351}, {32, 0.63047350420733980638},
            {64, 0.63047350337438679649}, {128, 0.63047350337438679612}, {256, 0.63047350337438679612},
            {512, 0.63047350337438679612}, {1024, 0.63047350337438679612}};
        auto it = t.find(m); return it == t.end() ? 0.0 : it->second;
    }
    if (a == 'd') {
        if (m == 2) return 1.0 / 3;
        if (m == 3) return 8.0 + tpi * tpi / 3 - 27.0 * c3;
        if (m == 4) return tpi * tpi / 9.0 - 1.0 / 6 - 7.0 * zeta3 / 2;
        if (m == 5) return 0.0024846057623403154800;
        if (m == 6) return 0.00048914170018803477510;
        return 0.0;
    }
    if (m == 2) return 6.0 - 8.0 * log2;
    if (m == 3) return 10.0 - tpi * tpi / 2 - 8.0 * log2 + 32.0 * log2 * log2;
    if (m == 4) return 22.0 - 82.0 * zeta3 - 24.0 * log2 + 176.0 * log2 * log2 - 256.0 * log2 * log2 * log2 / 3 + 4.0 * (tpi * tpi) * log2 - 11.0 * tpi * tpi / 6.0;
    if (m == 5) return 0.0034936537117295217407;
    if (m == 6) return 0.00068783287182640943700;
    return 0.0;
}

int main(int argc, char** argv) {
    char a = drv::arg_a(argc, argv, 1, 'c');
    int m = drv::arg_i(argc, argv, 2, 6), n = drv::arg_i(argc, argv, 3, 65), r = drv::arg_i(argc, argv, 4, 20), piv = drv::arg_i(argc, argv, 5, 1);
    int adj = 0;
    if (n % 2 == 0) { n += 1; adj = 1; }
    int nparts = std::getenv("TTC_PARTITIONS") ? std::atoi(std::getenv("TTC_PARTITIONS")) : 1;
    std::printf("Hi, this is TT cross interpolation computing Ising integral...\n");
    std::printf("   integral :%10c\n", a);
    std::printf("   dimension:%10d\n", m);
    drv::banner_common(n, adj, r, piv, nparts);
    const double acc = 500 * 2.220446049250313e-16;
    a = (char)std::tolower(a);
    std::vector<double> par(2 * n + 1);
    if (a == 'c') par[2 * n] = 1; else if (a == 'd') par[2 * n] = 2; else if (a == 'e') par[2 * n] = 3;
    else { std::printf(" unknown integral type:%c\n", a); return 1; }
    double tru = ising_tru(a, m);
    std::vector<double> x(n), w(n);
    ttc_lgwt(n, x.data(), w.data());
    for (int i = 0; i < n; ++i) { w[i] = 0.5 * w[i]; x[i] = (x[i] + 1.0) / 2; }
    const bool rescale = (a == 'd' || a == 'e') && m >= 10;
    const double val = (double)(n / 2);
    for (int i = 0; i < n; ++i) { w[i] = (rescale ? 5.0 * val : val) * w[i]; par[i] = x[i]; par[n + i] = w[i]; }
    const int d = m - 1;
    std::vector<int> nn(d, n);
    std::vector<double> quad((size_t)d * n, 1.0 / val);
    ttc_handle* h = nullptr;
    int st = ttc_create(&h, TTC_ISING, d, nn.data(), par.data(), (long)par.size(), nullptr, 0);
    if (st) drv::die(nullptr, st, "ttc_create");
    ttc_set_quad(h, quad.data());
    if (tru != 0.0) ttc_set_tru(h, 1, tru);
    int rc = drv::run_and_report(h, r, acc, piv, tru, tru != 0.0, rescale ? m - 1 : -1);
    ttc_destroy(h);
    return rc;
}
