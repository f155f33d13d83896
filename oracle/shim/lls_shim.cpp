// oracle/shim/lls_shim.cpp -- TEST INFRASTRUCTURE.  Replaces /root/reference/source/numeric/mylapack.cpp,
// whose only live function (Cmylapack::lls, lines 137-147) calls Eigen 3.3.7's jacobiSvd().solve() --
// Eigen is fetched by git (lib/CMakeLists.txt:13-17) and absent here.  Same contract: minimise |Ax-b|
// in double, return float.  Solved by Householder QR (full-rank, tiny n).  PARITY UNPINNED (Eigen).
#include <cmath>
#include <vector>

#include "numeric/mylapack.hpp"

void Cmylapack::lls(const std::vector<std::vector<float> >& A, const std::vector<float>& b, std::vector<float>& x) {
  const int m = (int)A.size();
  const int n = (int)A[0].size();
  std::vector<double> a((size_t)m * n), r(m);
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = A[i][j];
    r[i] = b[i];
  }
  for (int k = 0; k < n && k < m; ++k) {
    double nrm = 0.0;
    for (int i = k; i < m; ++i) nrm += a[(size_t)i * n + k] * a[(size_t)i * n + k];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    const double alpha = a[(size_t)k * n + k] > 0 ? -nrm : nrm;
    std::vector<double> v(m, 0.0);
    for (int i = k; i < m; ++i) v[i] = a[(size_t)i * n + k];
    v[k] -= alpha;
    double vn = 0.0;
    for (int i = k; i < m; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    for (int j = k; j < n; ++j) {
      double s = 0.0;
      for (int i = k; i < m; ++i) s += v[i] * a[(size_t)i * n + j];
      s = 2.0 * s / vn;
      for (int i = k; i < m; ++i) a[(size_t)i * n + j] -= s * v[i];
    }
    double s = 0.0;
    for (int i = k; i < m; ++i) s += v[i] * r[i];
    s = 2.0 * s / vn;
    for (int i = k; i < m; ++i) r[i] -= s * v[i];
  }
  std::vector<double> sol(n, 0.0);
  for (int k = n - 1; k >= 0; --k) {
    double s = r[k];
    for (int j = k + 1; j < n; ++j) s -= a[(size_t)k * n + j] * sol[j];
    const double d = a[(size_t)k * n + k];
    sol[k] = d != 0.0 ? s / d : 0.0;
  }
  x.resize(n);
  for (int i = 0; i < n; ++i) x[i] = (float)sol[i];
}
