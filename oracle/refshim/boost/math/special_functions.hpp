// oracle/refshim/boost/math/special_functions.hpp -- stand-in for the three Boost.Math
// entry points libcluster uses (digamma, lgamma, constants::pi); see Eigen/Dense in
// this directory for why this exists.  TEST INFRASTRUCTURE ONLY.
#ifndef LCB_REFSHIM_BOOST_MATH
#define LCB_REFSHIM_BOOST_MATH
#include <cmath>
namespace boost {
namespace math {
namespace constants {
template <typename T> inline T pi() { return (T)3.141592653589793238462643383279502884L; }
}  // namespace constants
// psi(x): reflection for x <= 0, upward recurrence to x >= 10, then the asymptotic series
template <typename T> inline T digamma(T xin) {
  long double x = (long double)xin, r = 0.0L;
  if (x <= 0.0L) {
    const long double pi = 3.141592653589793238462643383279502884L;
    return (T)((long double)digamma<long double>((long double)(1.0L - x)) - pi / std::tan(pi * x));
  }
  while (x < 10.0L) {
    r -= 1.0L / x;
    x += 1.0L;
  }
  const long double f = 1.0L / (x * x);
  const long double t = f * (-1.0L / 12 + f * (1.0L / 120 + f * (-1.0L / 252 + f * (1.0L / 240 + f * (-1.0L / 132 +
                        f * (691.0L / 32760 + f * (-1.0L / 12 + f * (3617.0L / 8160))))))));
  return (T)(r + std::log(x) - 0.5L / x + t);
}
template <typename T> inline T lgamma(T x) { return (T)std::lgamma((long double)x); }
}  // namespace math
}  // namespace boost
#endif
