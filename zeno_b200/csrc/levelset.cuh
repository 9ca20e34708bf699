// libflipb200 -- level-set fractions (FF/levelset_util.cpp:5-99), device versions.
#pragma once
#include <cuda_runtime.h>

namespace fb {
// fraction of the segment that is inside (phi < 0), FF/levelset_util.cpp:5-15
__device__ __forceinline__ float fraction_inside2(float l, float r) {
    if (l < 0.f && r < 0.f) return 1.f;
    if (l < 0.f && r >= 0.f) return __fdiv_rn(l, __fsub_rn(l, r));
    if (l >= 0.f && r < 0.f) return __fdiv_rn(r, __fsub_rn(r, l));
    return 0.f;
}
__device__ __forceinline__ void cycle4(float* a) {
    float t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = t;
}
// fraction of the unit square that is inside, FF/levelset_util.cpp:26-99
__device__ inline float fraction_inside4(float bl, float br, float tl, float tr) {
    int inside = (bl < 0.f ? 1 : 0) + (tl < 0.f ? 1 : 0) + (br < 0.f ? 1 : 0) + (tr < 0.f ? 1 : 0);
    float list[4] = {bl, br, tr, tl};
    if (inside == 4) return 1.f;
    if (inside == 3) {
        while (list[0] < 0.f) cycle4(list);
        float side0 = __fsub_rn(1.f, fraction_inside2(list[0], list[3]));
        float side1 = __fsub_rn(1.f, fraction_inside2(list[0], list[1]));
        return __fsub_rn(1.f, __fmul_rn(__fmul_rn(0.5f, side0), side1));
    }
    if (inside == 2) {
        while (list[0] >= 0.f || !(list[1] < 0.f || list[2] < 0.f)) cycle4(list);
        if (list[1] < 0.f) {
            float sl = fraction_inside2(list[0], list[3]);
            float sr = fraction_inside2(list[1], list[2]);
            return __fmul_rn(0.5f, __fadd_rn(sl, sr));
        }
        float middle = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(list[0], list[1]), list[2]), list[3]));
        if (middle < 0.f) {
            float area = 0.f;
            float side1 = __fsub_rn(1.f, fraction_inside2(list[0], list[3]));
            float side3 = __fsub_rn(1.f, fraction_inside2(list[2], list[3]));
            area = __fadd_rn(area, __fmul_rn(__fmul_rn(0.5f, side1), side3));
            float side2 = __fsub_rn(1.f, fraction_inside2(list[2], list[1]));
            float side0 = __fsub_rn(1.f, fraction_inside2(list[0], list[1]));
            area = __fadd_rn(area, __fmul_rn(__fmul_rn(0.5f, side0), side2));
            return __fsub_rn(1.f, area);
        }
        float area = 0.f;
        float side0 = fraction_inside2(list[0], list[1]);
        float side1 = fraction_inside2(list[0], list[3]);
        area = __fadd_rn(area, __fmul_rn(__fmul_rn(0.5f, side0), side1));
        float side2 = fraction_inside2(list[2], list[1]);
        float side3 = fraction_inside2(list[2], list[3]);
        area = __fadd_rn(area, __fmul_rn(__fmul_rn(0.5f, side2), side3));
        return area;
    }
    if (inside == 1) {
        while (list[0] >= 0.f) cycle4(list);
        float side0 = fraction_inside2(list[0], list[3]);
        float side1 = fraction_inside2(list[0], list[1]);
        return __fmul_rn(__fmul_rn(0.5f, side0), side1);
    }
    return 0.f;
}
}  // namespace fb
