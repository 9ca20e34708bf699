// TEST INFRASTRUCTURE ONLY. zeno::NumericObject as the plugin uses it (zeno/include/zeno/types/NumericObject.h:15-39).
#pragma once
#include <zeno/zeno.h>
namespace zeno {
struct NumericObject : IObject {
    std::variant<int, float, vec3f> value;
    NumericObject() : value(0) {}
    template <class T> explicit NumericObject(T v) : value(v) {}
    template <class T> T get() const {
        if (auto p = std::get_if<T>(&value)) return *p;
        if constexpr (std::is_same_v<T, float>) { if (auto q = std::get_if<int>(&value)) return float(*q); }
        throw std::runtime_error("NumericObject::get: wrong type");
    }
    template <class T> void set(T v) { value = v; }
};
template <class T> T INode::get_input2(std::string const& id) const {
    if (is_param(id)) return get_param<T>(id.substr(0, id.size() - 1));
    if constexpr (std::is_same_v<T, std::string>) throw std::runtime_error("INode::get_input2<string>: `" + id + "` is not a param");
    else return get_input<NumericObject>(id)->template get<T>();
}
}  // namespace zeno
