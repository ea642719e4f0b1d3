// Distribution functions used by the host-side AU fit.  They must evaluate the SAME
// published algorithms with the same constants and operation order as the reference
// (src/rand.c) so that printed p-values agree digit for digit:
//   pnorm  = 0.5*erfc(-x/sqrt2) with erfc from the incomplete gamma function
//            (series / continued fraction, Numerical Recipes 2nd ed. ch. 6.2)   src/rand.c:517-651
//   qnorm  = bisection on pnorm over [-37,37] to 1e-10                           src/rand.c:465-493
//   dnorm  = exp(-x^2/2)/sqrt(2 pi)                                              src/rand.c:455-458
//   poz    = Ibbetson, CACM Algorithm 209 (as adapted by G. Perlman)             src/rand.c:137-181
//   pochisq= Hill & Pike, CACM Algorithm 299 (Perlman)                           src/rand.c:244-295
#pragma once
namespace specfun {
double dnorm(double x);
double pnorm(double x);
double qnorm(double p);
void qnorm_set_fast(int on);  // 0: plain 40-step bisection at every call (tests compare the two)
bool qnorm_grid_is_monotone();  // the computed pnorm on the 2^20-cell grid of qnorm's first 20 halvings
double poz(double z);
double pochisq(double x, int df);
}  // namespace specfun
