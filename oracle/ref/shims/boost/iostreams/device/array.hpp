// Stand-in for <boost/iostreams/device/array.hpp>.
#pragma once
#include <cstddef>
namespace boost { namespace iostreams {
struct array_source { const char* begin; std::size_t size; };
}}
