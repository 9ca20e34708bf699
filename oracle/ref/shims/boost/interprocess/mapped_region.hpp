// Stand-in for <boost/interprocess/mapped_region.hpp>: mmap of the whole file.
#pragma once
#include "file_mapping.hpp"
#include <sys/mman.h>
#include <sys/stat.h>
namespace boost { namespace interprocess {
class mapped_region {
public:
    mapped_region(const file_mapping& m, mode_t) {
        struct stat st; if (::fstat(m.fd(), &st) != 0) throw std::runtime_error("mapped_region: fstat failed");
        mSize = (std::size_t)st.st_size;
        mAddr = mSize ? ::mmap(nullptr, mSize, PROT_READ, MAP_PRIVATE, m.fd(), 0) : nullptr;
        if (mAddr == MAP_FAILED) throw std::runtime_error("mapped_region: mmap failed");
    }
    ~mapped_region() { if (mAddr && mSize) ::munmap(mAddr, mSize); }
    void* get_address() const { return mAddr; }
    std::size_t get_size() const { return mSize; }
private:
    void* mAddr = nullptr; std::size_t mSize = 0;
};
}}
