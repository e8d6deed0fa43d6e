// common.cuh -- error plumbing, launch accounting and the counter-based RNG shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
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

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// Inside the lockstep loops every kernel depends on its predecessor, so each kernel boundary would cost a full
// drain + launch + prologue.  Kernels of the loop are launched with programmaticStreamSerialization: a CTA of
// kernel k+1 may start while kernel k is still running, does what does not depend on k (TMEM allocation, mbarrier
// init, TMA of weights / loads of state last written >= 2 kernels back) and then blocks in griddepcontrol.wait
// until k has completed and its writes are visible.  Every loop kernel triggers its dependents right after its own
// wait, so kernel k+1 only ever overlaps kernel k (everything <= k-1 is complete when k+1's prologue runs).
// Launched without the attribute, both instructions are no-ops.
extern std::atomic<int> g_pdl;            // uavrl_set_pdl(); default on

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// which loop kernel was launched last on the learner's stream (decides what a prologue may touch before its wait)
enum PdlPrev { kPdlNone = 0, kPdlAct, kPdlEnv, kPdlTd, kPdlTrain, kPdlDw, kPdlAdam };
// TcArgs.pdl / kernel flags
constexpr int kPdlOn = 1, kPdlEarlyWeights = 2, kPdlEarlyRows = 4;

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
