// Stand-in for <boost/uuid/uuid_io.hpp>: canonical 8-4-4-4-12 text form.
#pragma once
#include "uuid.hpp"
#include <istream>
#include <ostream>
#include <string>
namespace boost { namespace uuids {
inline std::string to_string(const uuid& u) {
    static const char* h = "0123456789abcdef";
    std::string s;
    for (int i = 0; i < 16; i++) { s += h[u.data[i] >> 4]; s += h[u.data[i] & 15]; if (i == 3 || i == 5 || i == 7 || i == 9) s += '-'; }
    return s;
}
inline std::ostream& operator<<(std::ostream& os, const uuid& u) { return os << to_string(u); }
inline std::istream& operator>>(std::istream& is, uuid& u) {
    auto hv = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1)); };
    for (int i = 0; i < 16; i++) {
        char a, b;
        if (!(is.get(a))) return is;
        if (a == '-') { if (!(is.get(a))) return is; }
        if (!(is.get(b))) return is;
        int x = hv(a), y = hv(b);
        if (x < 0 || y < 0) { is.setstate(std::ios::failbit); return is; }
        u.data[i] = (uint8_t)(x * 16 + y);
    }
    return is;
}
}}
