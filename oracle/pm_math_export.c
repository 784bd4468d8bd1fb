/* pm_math_export.c -- exports the portable libm of cilqr_b200/csrc/pm_math.h (static inline there) so that
 * tests/test_pm_math.py can measure it against glibc.  TEST INFRASTRUCTURE ONLY. */
#include "../cilqr_b200/csrc/pm_math.h"
double pm_export_sin(double x) { return pm_sin(x); }
double pm_export_cos(double x) { return pm_cos(x); }
double pm_export_tan(double x) { return pm_tan(x); }
double pm_export_log(double x) { return pm_log(x); }
double pm_export_hypot(double x, double y) { return pm_hypot(x, y); }
void pm_export_batch(int which, int n, const double* x, const double* y, double* out) {
  for (int i = 0; i < n; ++i)
    out[i] = which == 0 ? pm_sin(x[i]) : which == 1 ? pm_cos(x[i]) : which == 2 ? pm_tan(x[i]) : which == 3 ? pm_log(x[i]) : pm_hypot(x[i], y[i]);
}
/* the same five functions from the C library this oracle is linked against (glibc), for the comparison */
void pm_export_libm_batch(int which, int n, const double* x, const double* y, double* out) {
  for (int i = 0; i < n; ++i)
    out[i] = which == 0 ? sin(x[i]) : which == 1 ? cos(x[i]) : which == 2 ? tan(x[i]) : which == 3 ? log(x[i]) : hypot(x[i], y[i]);
}
