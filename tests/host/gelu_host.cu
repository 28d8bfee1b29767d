// Host-side check of the device GELU helpers (compiled by nvcc for the CPU only): sweeps gelu_erf_f / gelu_erf_grad_f of
// svd_xtend_b200/csrc/common.cuh against erf() in double precision and prints the error summary as one line of JSON.
#include <cmath>
#include <cstdio>
#include "../../svd_xtend_b200/csrc/common.cuh"

int main() {
  double max_abs = 0, max_rel = 0, max_abs_g = 0, max_rel_g = 0;
  long n = 0;
  for (double xd = -9.0; xd <= 9.0; xd += 1.0 / 4096.0, ++n) {
    const float x = (float)xd;
    const double xe = (double)x;
    const double cdf = 0.5 * (1.0 + erf(xe / sqrt(2.0)));
    const double ref = xe * cdf;
    const double refg = cdf + xe * exp(-0.5 * xe * xe) / sqrt(2.0 * M_PI);
    const double got = (double)svdx::gelu_erf_f(x), gotg = (double)svdx::gelu_erf_grad_f(x);
    const double ea = fabs(got - ref), eg = fabs(gotg - refg);
    if (ea > max_abs) max_abs = ea;
    if (eg > max_abs_g) max_abs_g = eg;
    if (fabs(xe) <= 5.5 && fabs(ref) > 1e-30 && ea / fabs(ref) > max_rel) max_rel = ea / fabs(ref);
    if (fabs(xe) <= 5.5 && fabs(refg) > 1e-3 && eg / fabs(refg) > max_rel_g) max_rel_g = eg / fabs(refg);
  }
  printf("{\"n\": %ld, \"max_abs\": %.3e, \"max_rel\": %.3e, \"grad_max_abs\": %.3e, \"grad_max_rel\": %.3e}\n", n, max_abs, max_rel,
         max_abs_g, max_rel_g);
  return 0;
}
