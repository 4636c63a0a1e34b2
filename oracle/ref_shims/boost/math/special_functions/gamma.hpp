// Stand-in for <boost/math/special_functions/gamma.hpp> (Boost 1.84 is not present under
// /root/reference).  TEST INFRASTRUCTURE ONLY.  gamma_q / gamma_q_inv forward to the oracle's fp64
// restatement in oracle/gamma_q.h (values pinned on a 40-digit mpmath table, not on Boost's binary; see that header).
#ifndef DPHY_ORACLE_SHIM_BOOST_MATH_GAMMA_HPP_
#define DPHY_ORACLE_SHIM_BOOST_MATH_GAMMA_HPP_
#include "gamma_q.h"
namespace boost { namespace math {
template<typename Policy> inline double gamma_q(double a, double x, const Policy&) { return orc_gamma_q(a, x); }
inline double gamma_q(double a, double x) { return orc_gamma_q(a, x); }
template<typename Policy> inline double gamma_q_inv(double a, double q, const Policy&) { return orc_gamma_q_inv(a, q); }
inline double gamma_q_inv(double a, double q) { return orc_gamma_q_inv(a, q); }
}}
#endif
