// TEST-ONLY: exposes option-pricing-ffn-lbfgs_b200/csrc/dhj_fastmath.cuh to ctypes for accuracy checks on the CPU.
#include "dhj_fastmath.cuh"
using namespace dhj::fm;
extern "C" {
void fm_sincos(const double* x, int n, double* s, double* c) { for (int i = 0; i < n; ++i) sincos_(x[i], s + i, c + i); }
void fm_exp(const double* x, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = exp_(x[i]); }
void fm_exp_tab(const double* x, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = exp_tab(x[i], &kTables); }
void fm_log_ratio(const double* a, const double* b, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = log_ratio(a[i], b[i]); }
void fm_log_tab(const double* a, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = log_tab(a[i], &kTables); }
void fm_atan2(const double* y, const double* x, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = atan2_(y[i], x[i]); }
void fm_atan2_tab(const double* y, const double* x, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = atan2_tab(y[i], x[i], &kTables); }
void fm_div(const double* a, const double* b, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = div(a[i], b[i]); }
void fm_rcp(const double* a, int n, double* o) { for (int i = 0; i < n; ++i) o[i] = rcp(a[i]); }
void fm_sqrt(const double* a, int n, double* s, double* y) { for (int i = 0; i < n; ++i) sqrt_rsqrt(a[i], s + i, y + i); }
}
