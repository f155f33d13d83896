// oracle/shim/nlopt.hpp -- TEST INFRASTRUCTURE.  Stand-in for nlopt 2.6.1's C++ header.
//
// The reference's optimiser call site (/root/reference/source/pmvs/optim.cpp:621-644) uses
// nlopt::opt(LN_BOBYQA, 3) with set_min_objective / set_xtol_rel / set_maxeval /
// set_lower_bounds / set_upper_bounds / optimize and nlopt::srand.  nlopt itself is fetched by
// git at configure time (lib/CMakeLists.txt:19-25) and is absent here, so this header supplies
// exactly that API subset on top of oracle/nm3.h (the Nelder-Mead the whole repo uses).
// Whatever algorithm id is requested, the minimiser that runs is nm3.  PARITY UNPINNED.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "../nm3.h"

// Stopping tolerance of the stand-in in scaled-parameter units (1 = half a pixel of image motion in depth, pi/48 in the
// angles).  The reference asks nlopt for xtol_rel 1e-7; a simplex run down to that walks a plateau of the f32 objective
// (178 instead of 104 evaluations per patch) and ends where the 1e-3 run ends: against a restatement of BOBYQA at the
// reference's settings the three tolerances 1e-3 / 1e-4 / 1e-7 give the same percentiles to four digits
// (profiles/r2_optimiser_bound_small.json; tools/research/optimiser_bound_study.py).  Runtime-settable (xtol_floor()).
#ifndef PMVS_NM_XTOL_FLOOR
#define PMVS_NM_XTOL_FLOOR 1.0e-3
#endif
#ifndef PMVS_NM_STEP
#define PMVS_NM_STEP 1.0
#endif

namespace nlopt {
enum algorithm { LN_NELDERMEAD, LN_SBPLX, LN_COBYLA, LN_BOBYQA, LN_PRAXIS };
enum result {
  FAILURE = -1, INVALID_ARGS = -2, OUT_OF_MEMORY = -3, ROUNDOFF_LIMITED = -4, FORCED_STOP = -5,
  SUCCESS = 1, STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4, MAXEVAL_REACHED = 5,
  MAXTIME_REACHED = 6
};
typedef double (*func)(unsigned n, const double* x, double* grad, void* f_data);

inline void srand(unsigned long) {}

// Counters the probe driver reads (objective evaluations across all optimize() calls).
inline unsigned long long& total_evals() { static unsigned long long v = 0; return v; }
inline unsigned long long& total_calls() { static unsigned long long v = 0; return v; }
inline int& last_evals() { static thread_local int v = 0; return v; }
// The x-tolerance floor at run time (default PMVS_NM_XTOL_FLOOR; 0 = exactly the reference's xtol_rel * step = 1e-7):
// the optimiser-tolerance studies and bench.py's second CPU figure set it through oracle/ref_probe.cpp.
inline double& xtol_floor() { static double v = PMVS_NM_XTOL_FLOOR; return v; }

class opt {
 public:
  opt(algorithm, unsigned n) : _n(n), _f(nullptr), _data(nullptr), _xtol_rel(0.0), _maxeval(0),
                               _lb(n, -HUGE_VAL), _ub(n, HUGE_VAL) {
    if (n > NM3_MAXN) throw std::invalid_argument("nm3 shim: n too large");
  }
  void set_min_objective(func f, void* data) { _f = f; _data = data; }
  void set_xtol_rel(double t) { _xtol_rel = t; }
  void set_maxeval(int m) { _maxeval = m; }
  void set_lower_bounds(const std::vector<double>& lb) { _lb = lb; }
  void set_upper_bounds(const std::vector<double>& ub) { _ub = ub; }

  result optimize(std::vector<double>& x, double& minf) {
    if (!_f || x.size() != _n) throw std::invalid_argument("nm3 shim: bad arguments");
    for (unsigned i = 0; i < _n; ++i)
      if (x[i] < _lb[i] || x[i] > _ub[i]) throw std::invalid_argument("nm3 shim: x out of bounds");
    const double step = PMVS_NM_STEP;
    const double xtol = std::fmax(_xtol_rel * step, xtol_floor());
    int nev = 0;
    const int maxeval = _maxeval > 0 ? _maxeval : 1000000;
    int r = nm3_minimize(_n, &opt::thunk, this, _lb.data(), _ub.data(), x.data(), &minf, step,
                         xtol, maxeval, &nev);
    last_evals() = nev;
    __atomic_fetch_add(&total_evals(), (unsigned long long)nev, __ATOMIC_RELAXED);
    __atomic_fetch_add(&total_calls(), 1ULL, __ATOMIC_RELAXED);
    return (result)r;
  }

 private:
  static double thunk(unsigned n, const double* x, void* self) {
    opt* o = (opt*)self;
    return o->_f(n, x, nullptr, o->_data);
  }
  unsigned _n;
  func _f;
  void* _data;
  double _xtol_rel;
  int _maxeval;
  std::vector<double> _lb, _ub;
};
}  // namespace nlopt
