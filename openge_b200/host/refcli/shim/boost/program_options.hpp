// Stand-in for <boost/program_options.hpp> (Boost is not installed in this image).
// The reference's algorithm layer includes commands/commands.h, which names these
// types as members but never uses them on the dedup path. Test infrastructure only.
#ifndef OGE_ORACLE_BOOST_PO_STUB_HPP
#define OGE_ORACLE_BOOST_PO_STUB_HPP
#include <string>
#include <set>
#include <map>
#include <vector>
#include <sstream>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <stdint.h>
namespace boost { namespace program_options {
class positional_options_description {};
class options_description {};
class variables_map {};
} }
#endif
