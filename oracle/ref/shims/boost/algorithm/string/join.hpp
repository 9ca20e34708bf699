// Stand-in for <boost/algorithm/string/join.hpp> (openvdb/io/Compression.cc:31).
#pragma once
#include <string>
namespace boost {
template <typename Seq> std::string join(const Seq& words, const std::string& sep) {
    std::string out; bool first = true;
    for (const auto& w : words) { if (!first) out += sep; out += w; first = false; }
    return out;
}
}
