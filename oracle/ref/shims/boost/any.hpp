// Stand-in for <boost/any.hpp>: std::any (openvdb/io/io.h:92, io/Archive.cc:327-330).
#pragma once
#include <any>
namespace boost {
using any = std::any;
using bad_any_cast = std::bad_any_cast;
template <typename T, typename A> T any_cast(A&& a) { return std::any_cast<T>(std::forward<A>(a)); }
}
