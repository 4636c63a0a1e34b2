// tools/delphy.cpp:4,14 includes this header and aliases the namespace but uses nothing from it (flags go through cxxopts).
#ifndef DPHY_SHIM_BOOST_PROGRAM_OPTIONS_
#define DPHY_SHIM_BOOST_PROGRAM_OPTIONS_
namespace boost { namespace program_options {} }
#endif
