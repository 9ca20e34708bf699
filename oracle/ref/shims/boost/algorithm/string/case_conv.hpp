// Stand-in for <boost/algorithm/string/case_conv.hpp> (openvdb/Grid.cc:128, math/FiniteDifference.h:84).
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
namespace boost {
inline void to_lower(std::string& s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); }); }
inline void to_upper(std::string& s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::toupper(c); }); }
}
