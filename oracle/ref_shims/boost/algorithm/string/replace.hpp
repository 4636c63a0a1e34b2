// Shim for boost::replace_all / replace_all_copy (core/io.cpp:265, core/newick.cpp:256); Boost is not in this image.
// BUILD INFRASTRUCTURE for compiling the reference in place -- never shipped.
#ifndef DPHY_SHIM_BOOST_REPLACE_
#define DPHY_SHIM_BOOST_REPLACE_
#include <string>
namespace boost {
inline void replace_all(std::string& s, const std::string& from, const std::string& to) {
  if (from.empty()) { return; }
  for (std::string::size_type pos = 0; (pos = s.find(from, pos)) != std::string::npos; pos += to.size()) {
    s.replace(pos, from.size(), to);
  }
}
inline std::string replace_all_copy(std::string s, const std::string& from, const std::string& to) {
  replace_all(s, from, to);
  return s;
}
}  // namespace boost
#endif
