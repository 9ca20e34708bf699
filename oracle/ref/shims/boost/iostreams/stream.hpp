// Stand-in for <boost/iostreams/stream.hpp>: a read-only streambuf over a memory range
// (openvdb/io/Archive.cc:559).
#pragma once
#include "device/array.hpp"
#include <streambuf>
namespace boost { namespace iostreams {
template <typename Device> class stream_buffer;
template <> class stream_buffer<array_source> : public std::streambuf {
public:
    stream_buffer(const char* p, std::size_t n) { char* b = const_cast<char*>(p); setg(b, b, b + n); }
protected:
    pos_type seekoff(off_type off, std::ios_base::seekdir dir, std::ios_base::openmode) override {
        char* base = eback(); char* cur = gptr(); char* end = egptr();
        char* t = dir == std::ios_base::beg ? base + off : (dir == std::ios_base::cur ? cur + off : end + off);
        if (t < base || t > end) return pos_type(off_type(-1));
        setg(base, t, end); return pos_type(t - base);
    }
    pos_type seekpos(pos_type pos, std::ios_base::openmode m) override { return seekoff(off_type(pos), std::ios_base::beg, m); }
};
}}
