// common.cuh -- error plumbing, launch accounting and the counter-based RNG shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/uavrl.h"

namespace uavrl {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

#define UAVRL_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            char _b[512];                                                                     \
            snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,      \
                     cudaGetErrorString(_e));                                                 \
            return ::uavrl::fail(UAVRL_ERR_CUDA, _b);                                         \
        }                                                                                     \
    } while (0)

#define UAVRL_LAUNCHED()                                                                      \
    do {                                                                                      \
        ::uavrl::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
        UAVRL_CUDA(cudaGetLastError());                                                       \
    } while (0)

template <class T>
inline int dev_alloc(T **p, size_t n)
{
    UAVRL_CUDA(cudaMalloc((void **)p, n * sizeof(T)));
    UAVRL_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: stream = (key, counter), no state to store.
struct Philox {
    static __host__ __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
    {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0;
        const uint32_t h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
        const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
    }
    static __host__ __device__ __forceinline__ void gen(uint64_t key, uint64_t ctr_lo, uint64_t ctr_hi,
                                                        uint32_t (&out)[4])
    {
        uint32_t c[4] = { (uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), (uint32_t)ctr_hi, (uint32_t)(ctr_hi >> 32) };
        uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
    // uniform in [0,1) with 24 random bits (what a float can hold exactly)
    static __host__ __device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
};

}  // namespace uavrl
