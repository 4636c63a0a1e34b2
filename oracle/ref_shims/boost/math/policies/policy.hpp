// Stand-in for <boost/math/policies/policy.hpp> (Boost 1.84 is not present under /root/reference).
// TEST INFRASTRUCTURE ONLY -- lets core/safe_gamma_math.h compile for oracle/_ref.
#ifndef DPHY_ORACLE_SHIM_BOOST_MATH_POLICY_HPP_
#define DPHY_ORACLE_SHIM_BOOST_MATH_POLICY_HPP_
namespace boost { namespace math { namespace policies {
template<bool B> struct promote_double {};
template<typename... Ts> struct policy {};
}}}
#endif
