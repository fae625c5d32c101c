// Per-device kernel configuration cache.
//
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the occupancy query are per DEVICE, and the C ABI lets one
// host thread drive several devices (sr_init(device)), so the cache is an array indexed by the current device,
// one per kernel instantiation (the struct is a function-local static of the templated launcher).
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace sr {

struct KernelCache {
    static constexpr int MAX_DEVICES = 64;
    std::atomic<int> blocks[MAX_DEVICES];  // resident CTAs per SM; 0 = not configured on that device yet
    KernelCache() {
        for (int i = 0; i < MAX_DEVICES; i++) blocks[i].store(0, std::memory_order_relaxed);
    }
    // Sets the dynamic shared-memory limit of `kern` on the current device (first call per device) and returns the
    // number of CTAs of `threads` threads / `smem` bytes that fit on one SM (>= 1).
    template <class K>
    cudaError_t configure(K kern, int threads, size_t smem, int* blocks_per_sm) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
        int b = blocks[dev].load(std::memory_order_acquire);
        if (b == 0) {  // (idempotent: two threads racing here both set the same attribute)
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, threads, smem);
            if (e != cudaSuccess) return e;
            if (b < 1) b = 1;
            blocks[dev].store(b, std::memory_order_release);
        }
        *blocks_per_sm = b;
        return cudaSuccess;
    }
};

}  // namespace sr
