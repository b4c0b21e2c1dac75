// What a traversal kernel needs to know about the flattened BVH (bvh.h). Passed to kernels by value, so it lives in
// the constant bank.
#pragma once
#include <cuda_runtime.h>

namespace lmb200 {

struct BvhDev {
    const float4* units;     // 64-byte units: nodes and triangle units, root = unit 0
    float gstep[3];          // scene grid step per axis (power of two)
    float glo2[3];           // scene grid origin - 2^23 * step: node origin = fma(2^23 + k, step, glo2), exact (bvh.h)
};

}  // namespace lmb200
