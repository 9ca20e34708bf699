// Stand-in for <boost/uuid/uuid.hpp>: 16 raw bytes (openvdb/io/Archive.h:183).
#pragma once
#include <cstdint>
#include <cstring>
#include <algorithm>
namespace boost { namespace uuids {
struct uuid {
    uint8_t data[16];
    using iterator = uint8_t*; using const_iterator = const uint8_t*;
    iterator begin() { return data; } iterator end() { return data + 16; }
    const_iterator begin() const { return data; } const_iterator end() const { return data + 16; }
    bool is_nil() const { for (auto b : data) if (b) return false; return true; }
    static constexpr std::size_t static_size() { return 16; }
    std::size_t size() const { return 16; }
};
inline bool operator==(const uuid& a, const uuid& b) { return std::memcmp(a.data, b.data, 16) == 0; }
inline bool operator!=(const uuid& a, const uuid& b) { return !(a == b); }
inline bool operator<(const uuid& a, const uuid& b) { return std::memcmp(a.data, b.data, 16) < 0; }
}}
