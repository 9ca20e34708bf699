// TEST INFRASTRUCTURE ONLY -- a seeded stand-in for std::random_device, force-included when oracle/ref/build_ref.sh compiles the
// reference's FLIP_vdb.cpp (never edited: the header arrives through `-include random -include seeded_random.h`).
// FLIP_vdb::reseed_fluid / emit_liquid draw where their jitter table starts from `std::random_device device; std::mt19937
// gen(device());` (FF/FLIP_vdb.cpp:2081-2084,2358-2361,2497-2500), once per TBB chunk; with this header `device()` returns
// flipref::seed(), so a run of the reference is reproducible and the oracle can be pinned against it.
#pragma once
#include <random>
namespace flipref {
inline unsigned& seed() { static unsigned s = 0u; return s; }
struct seeded_device {
    using result_type = unsigned;
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() { return seed(); }
};
}  // namespace flipref
namespace std { using flipref_seeded_device = ::flipref::seeded_device; }
#define random_device flipref_seeded_device
