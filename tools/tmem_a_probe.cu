// Probe for round 2: tcgen05.mma with the A operand in TMEM (kind::tf32).  (1) correctness of the assumed layout
// (lane = row m, one 32-bit column per k, 8 columns per instruction) against a CPU product; (2) cycles per
// M=128 x N=64 x K=8 instruction with A in SMEM (both operands read from shared memory) vs A in TMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tmem_a_probe tools/tmem_a_probe.cu && /tmp/tmem_a_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../dqn-based-uav-3d_path_planer_b200/csrc/tma.cuh"
#include "../dqn-based-uav-3d_path_planer_b200/csrc/umma.cuh"
using namespace uavrl;

constexpr int N = 64, K = 64;

__device__ __forceinline__ void umma_tf32_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

// mode 0: A in SMEM, mode 1: A in TMEM.  reps: how often the K loop is re-issued (timing)
__global__ void __launch_bounds__(128) probe(const float *A, const float *B, float *D, long long *cyc, int mode, int reps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t SBO = umma_sbo(K);
    unsigned char *As = smem, *Bs = smem + umma_tile_bytes(128, K);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, 128);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, c = i % K; *reinterpret_cast<float *>(As + umma_off(r, c, SBO)) = A[r * K + c]; }
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, c = i % K; *reinterpret_cast<float *>(Bs + umma_off(r, c, SBO)) = B[r * K + c]; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base, ta = tb + 64;                   // D: columns 0..63, A: columns 64..127
    if (mode == 1) {                                               // this thread's row -> its TMEM lane, K columns
        for (int c0 = 0; c0 < K; c0 += 8) {
            float v[8];
            for (int j = 0; j < 8; ++j) v[j] = A[tid * K + c0 + j];
            tmem_st8(ta + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = 0;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_tf32(128, N);
        // descriptors precomputed, K loop fully unrolled: the loop body is the bare instruction stream
        uint64_t da[K / 8], db[K / 8]; uint32_t dta[K / 8];
#pragma unroll
        for (int k = 0; k < K / 8; ++k) {
            da[k] = umma_desc(smem_u32(As) + k * 2 * kUmmaLBO, SBO); db[k] = umma_desc(smem_u32(Bs) + k * 2 * kUmmaLBO, SBO);
            dta[k] = ta + (uint32_t)(8 * k);
        }
        t0 = clock64();
        if (mode == 0) {
            umma_tf32(tb, da[0], db[0], idesc, 0u);
#pragma unroll
            for (int k = 1; k < K / 8; ++k) umma_tf32(tb, da[k], db[k], idesc, 1u);
            for (int r = 1; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < K / 8; ++k) umma_tf32(tb, da[k], db[k], idesc, 1u);
            }
        } else {
            umma_tf32_ta(tb, dta[0], db[0], idesc, 0u);
#pragma unroll
            for (int k = 1; k < K / 8; ++k) umma_tf32_ta(tb, dta[k], db[k], idesc, 1u);
            for (int r = 1; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < K / 8; ++k) umma_tf32_ta(tb, dta[k], db[k], idesc, 1u);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) cyc[mode] = clock64() - t0;
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 128);
}

int main()
{
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(3);
    for (auto &x : A) x = (float)(rand() % 17 - 8);               // small integers: exact in TF32, exact sums
    for (auto &x : B) x = (float)(rand() % 9 - 4);
    float *dA, *dB, *dD; long long *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 16);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = umma_tile_bytes(128, K) + umma_tile_bytes(N, K);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode) {
        for (int reps : { 1, 16, 64 }) {
            cudaMemset(dD, 0, D.size() * 4);
            probe<<<1, 128, smem>>>(dA, dB, dD, dC, mode, reps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: CUDA error: %s\n", mode, cudaGetErrorString(e)); return 2; }
            long long c[2];
            cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    double s = 0;
                    for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * (double)B[n * K + k];
                    if (fabs(s * reps - D[r * N + n]) > 1e-3) ++bad;
                }
            printf("A in %s, %3d MMAs (M128 N64 K8): %lld cycles issue->commit = %.1f per MMA, mismatches vs CPU: %d / %d  (D[3][5]=%g)\n",
                   mode ? "TMEM" : "SMEM", reps * K / 8, c[mode], (double)c[mode] / (reps * K / 8), bad, 128 * N, D[3 * N + 5]);
        }
    }
    return 0;
}
