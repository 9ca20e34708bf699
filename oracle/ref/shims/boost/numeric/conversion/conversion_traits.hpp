// Stand-in for <boost/numeric/conversion/conversion_traits.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the reference's vendored OpenVDB 9.0.1 compile for oracle/_ref.
// openvdb/math/Math.h:925 uses only conversion_traits<S,T>::supertype.
#pragma once
#include <type_traits>
namespace boost { namespace numeric {
template <typename S, typename T, bool Arith = std::is_arithmetic<S>::value && std::is_arithmetic<T>::value>
struct conversion_traits { using supertype = S; };
template <typename S, typename T>
struct conversion_traits<S, T, true> {
    // Boost: the type with the larger range; float beats integer, wider beats narrower.
    using supertype = typename std::conditional<
        (std::is_floating_point<S>::value && !std::is_floating_point<T>::value), S,
        typename std::conditional<
            (std::is_floating_point<T>::value && !std::is_floating_point<S>::value), T,
            typename std::conditional<(sizeof(T) > sizeof(S)), T, S>::type>::type>::type;
};
}}
