// Small device helpers shared by the v2e_b200 translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ int warp_reduce_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_reduce_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// cudaFuncSetAttribute (the opt-in to more than 48 KB of dynamic shared memory) is per DEVICE: a process-wide flag
// would leave the second GPU of a process without it. One bit per device ordinal, set atomically.
#include <atomic>
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask[2];
    PerDeviceOnce() { mask[0] = 0; mask[1] = 0; }
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 127;
        const unsigned long long bit = 1ull << (dev & 63);
        return (mask[dev >> 6].fetch_or(bit) & bit) == 0;
    }
};
