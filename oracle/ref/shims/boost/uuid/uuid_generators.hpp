// Stand-in for <boost/uuid/uuid_generators.hpp> (openvdb/io/Archive.cc:601,1045).
#pragma once
#include "uuid.hpp"
namespace boost { namespace uuids {
inline uuid nil_uuid() { uuid u; std::memset(u.data, 0, 16); return u; }
template <typename Rng> struct basic_random_generator {
    Rng* rng;
    explicit basic_random_generator(Rng* r) : rng(r) {}
    uuid operator()() {
        uuid u;
        for (int i = 0; i < 16; i += 4) { uint32_t v = (uint32_t)(*rng)(); std::memcpy(u.data + i, &v, 4); }
        u.data[6] = (uint8_t)((u.data[6] & 0x0f) | 0x40);
        u.data[8] = (uint8_t)((u.data[8] & 0x3f) | 0x80);
        return u;
    }
};
}}
