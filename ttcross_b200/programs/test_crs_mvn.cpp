// C++ twin of test_crs_mvn.f90 (CLI: DIM N RANK PIV); setup follows test_crs_mvn.f90:23-133 and lib/mvn_pdf.f90:15-111.
#include "driver_common.hpp"

// Gauss-Jordan with partial pivoting (the reference calls LAPACK dgetrf/dgetri; any correct inverse is input data)
static void inv_det(std::vector<double> a, int n, std::vector<double>& inv, double& det) {
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
static double powi(double x, int m) {   // libgcc __powidf2 (what gfortran emits for real**integer)
    unsigned n = m < 0 ? -m : m;
    double y = (n % 2) ? x : 1.0;
    while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
    return m < 0 ? 1.0 / y : y;
}

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
    // mvn_init(d, 0, 1)
    const double sigma = 0.4, corr = 0.5, T = 1.0, rr = 0.0;
    std::vector<double> cov((size_t)d * d), inv;
    for (int i = 0; i < d; ++i) for (int j = 0; j < d; ++j) cov[(size_t)i * d + j] = ((i == j) ? sigma * sigma : sigma * corr * sigma) * T;
    double det;
    inv_det(cov, d, inv, det);
    std::vector<double> aux(d, std::log(100.0) + (rr - 0.5 * (sigma * sigma)) * T);
    aux.insert(aux.end(), inv.begin(), inv.end());
    aux.push_back(std::sqrt(powi(2.0 * 3.141592653589793, d) * det));
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
