// Stand-in for <boost/algorithm/string/predicate.hpp> (openvdb/io/GridDescriptor.cc:82).
#pragma once
#include <string>
namespace boost {
inline bool ends_with(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }
inline bool starts_with(const std::string& s, const std::string& pre) { return s.size() >= pre.size() && s.compare(0, pre.size(), pre) == 0; }
}
