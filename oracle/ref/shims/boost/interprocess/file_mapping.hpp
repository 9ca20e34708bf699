// Stand-in for <boost/interprocess/file_mapping.hpp>: POSIX open (openvdb/io/Archive.cc:455-514; delayed
// loading of .vdb files is not exercised by oracle/_ref).
#pragma once
#include <fcntl.h>
#include <unistd.h>
#include <cstdio>
#include <stdexcept>
#include <string>
namespace boost { namespace interprocess {
enum mode_t { read_only, read_write };
class file_mapping {
public:
    file_mapping(const char* name, mode_t) : mName(name) {
        mFd = ::open(name, O_RDONLY);
        if (mFd < 0) throw std::runtime_error(std::string("file_mapping: cannot open ") + name);
    }
    ~file_mapping() { if (mFd >= 0) ::close(mFd); }
    const char* get_name() const { return mName.c_str(); }
    int fd() const { return mFd; }
    static bool remove(const char* name) { return std::remove(name) == 0; }
private:
    std::string mName; int mFd = -1;
};
}}
