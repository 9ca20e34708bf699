// Stand-in for <boost/algorithm/string/erase.hpp> (openvdb/io/GridDescriptor.cc:84).
#pragma once
#include <string>
namespace boost {
inline void erase_last(std::string& s, const std::string& what) { size_t p = s.rfind(what); if (p != std::string::npos) s.erase(p, what.size()); }
}
