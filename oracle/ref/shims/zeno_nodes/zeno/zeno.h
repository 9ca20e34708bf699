// TEST INFRASTRUCTURE ONLY. A minimal stand-in for the part of the Zeno node runtime that the drop-in
// (zeno_b200/plugin/flipb200_nodes.cpp) uses: IObject / INode / NumericObject-free core, defNodeClass with the
// reference's Descriptor shape (zeno/include/zeno/core/Descriptor.h:9-45, INode.h:113-117 "param:" sockets) and makeError.
// It lets oracle/ref/plugin_nodes_test.cpp instantiate the plugin's node classes, wire real OpenVDB objects to their
// sockets and run apply() without building libzeno. Semantics kept: inputs are looked up by socket name, params are
// sockets named "<param>:", a missing input throws, defNodeClass registers a factory under the node's name.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>
#include <array>
#include <cstdio>
// IObject is the reference's own (header-only under ZENO_APIFREE): packed_FloatGrid3 derives from it and crosses into the
// FLIP_vdb statics compiled in other translation units, so its layout must be the real one
#ifndef ZENO_APIFREE
#define ZENO_APIFREE
#endif
#include <zeno/core/IObject.h>

namespace zeno {

struct vec3f : std::array<float, 3> {
    vec3f() : std::array<float, 3>{0.f, 0.f, 0.f} {}
    vec3f(float a, float b, float c) : std::array<float, 3>{a, b, c} {}
};

inline std::runtime_error makeError(const std::string& what) { return std::runtime_error(what); }

struct ParamDescriptor {
    std::string type, name, defl;
    ParamDescriptor(std::string const& t, std::string const& n, std::string const& d) : type(t), name(n), defl(d) {}
};
struct SocketDescriptor {
    std::string type, name, defl;
    SocketDescriptor(std::string const& t, std::string const& n, std::string const& d = {}) : type(t), name(n), defl(d) {}
    SocketDescriptor(const char* n) : SocketDescriptor({}, n) {}
};
struct Descriptor {
    std::vector<SocketDescriptor> inputs, outputs;
    std::vector<ParamDescriptor> params;
    std::vector<std::string> categories;
    Descriptor() = default;
    Descriptor(std::vector<SocketDescriptor> const& i, std::vector<SocketDescriptor> const& o, std::vector<ParamDescriptor> const& p,
               std::vector<std::string> const& c) : inputs(i), outputs(o), params(p), categories(c) {}
};

using ParamValue = std::variant<int, float, std::string>;

struct INode {
    std::map<std::string, std::shared_ptr<IObject>> inputs, outputs;
    std::map<std::string, ParamValue> params;
    virtual ~INode() = default;
    virtual void apply() = 0;
    // a param is the input socket "<name>:" (zeno/include/zeno/core/INode.h:107-111)
    static bool is_param(std::string const& id) { return !id.empty() && id.back() == ':'; }
    bool has_input(std::string const& id) const {
        return is_param(id) ? params.count(id.substr(0, id.size() - 1)) != 0 : inputs.count(id) != 0;
    }
    template <class T> bool has_input(std::string const& id) const {   // zeno/include/zeno/core/INode.h:84-89
        if (!has_input(id)) return false;
        auto it = inputs.find(id);
        return it != inputs.end() && dynamic_cast<T*>(it->second.get()) != nullptr;
    }
    std::shared_ptr<IObject> get_input(std::string const& id) const {
        auto it = inputs.find(id);
        if (it == inputs.end() || !it->second) throw std::runtime_error("INode::get_input: socket `" + id + "` is not connected");
        return it->second;
    }
    template <class T> std::shared_ptr<T> get_input(std::string const& id) const {
        auto p = std::dynamic_pointer_cast<T>(get_input(id));
        if (!p) throw std::runtime_error("INode::get_input<T>: socket `" + id + "` holds another object type");
        return p;
    }
    template <class T> T get_param(std::string const& id) const {
        auto it = params.find(id);
        if (it == params.end()) throw std::runtime_error("INode::get_param: no param `" + id + "`");
        if (auto p = std::get_if<T>(&it->second)) return *p;
        if constexpr (std::is_same_v<T, float>) { if (auto q = std::get_if<int>(&it->second)) return float(*q); }
        throw std::runtime_error("INode::get_param: param `" + id + "` has another type");
    }
    template <class T> T get_input2(std::string const& id) const;   // literal of a NumericObject socket or of a param (NumericObject.h)
    void set_output(std::string const& id, std::shared_ptr<IObject> obj) { outputs[id] = std::move(obj); }
};

struct NodeClass { std::function<std::unique_ptr<INode>()> make; Descriptor desc; };
inline std::map<std::string, NodeClass>& nodeRegistry() { static std::map<std::string, NodeClass> r; return r; }

template <class T>
int defNodeClass(std::string const& name, Descriptor const& desc = {}) {
    nodeRegistry()[name] = NodeClass{[] { return std::unique_ptr<INode>(new T()); }, desc};
    return 1;
}

}  // namespace zeno

// zeno/include/zeno/core/defNode.h:28-34
#define ZENDEFNODE(Class, ...) static int def##Class = zeno::defNodeClass<Class>(#Class, __VA_ARGS__)
#define ZENO_DEFNODE(Class) \
    static struct _Def##Class { _Def##Class(::zeno::Descriptor const& desc) { zeno::defNodeClass<Class>(#Class, desc); } } _def##Class
