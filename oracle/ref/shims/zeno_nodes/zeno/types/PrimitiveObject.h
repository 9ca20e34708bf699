// TEST INFRASTRUCTURE ONLY. zeno::PrimitiveObject as the drop-in's VDBPointsToPrimitive uses it
// (zeno/include/zeno/types/PrimitiveObject.h: resize, add_attr<vec3f>, attr<vec3f>, size).
#pragma once
#include <zeno/zeno.h>
#include <map>
#include <string>
#include <vector>
namespace zeno {
struct PrimitiveObject : IObject {
    std::map<std::string, std::vector<vec3f>> attrs;
    size_t n = 0;
    void resize(size_t count) { n = count; for (auto& kv : attrs) kv.second.resize(count); }
    size_t size() const { return n; }
    template <class T> std::vector<T>& add_attr(std::string const& name) { auto& a = attrs[name]; a.resize(n); return a; }
    template <class T> std::vector<T>& attr(std::string const& name) { return attrs.at(name); }
    bool has_attr(std::string const& name) const { return attrs.count(name) != 0; }
};
}  // namespace zeno
