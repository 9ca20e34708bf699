// TEST INFRASTRUCTURE ONLY -- stand-in for zeno/types/ConditionObject.h (FF/nosys/ParticleEmitter.cpp only asks whether a socket holds one)
#pragma once
#include <zeno/zeno.h>
namespace zeno {
struct ConditionObject : IObject {
    bool value;
    ConditionObject(bool value = true) : value(value) {}
    bool get() const { return value; }
};
}  // namespace zeno
