// TEST INFRASTRUCTURE ONLY. zeno::StringObject as the zenvdb nodes use it.
#pragma once
#include <zeno/zeno.h>
namespace zeno {
struct StringObject : IObject {
    std::string value;
    std::string const& get() const { return value; }
    void set(std::string const& v) { value = v; }
};
}  // namespace zeno
