// Standalone bring-up test of the tcgen05 building blocks (csrc/umma.cuh): one 128 x N x K GEMM tile,
// D = A * B^T with the 3xTF32 split, accumulator in TMEM, checked against an fp64 CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_gemm_test tools/tc_gemm_test.cu && ./tc_gemm_test
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../dqn-based-uav-3d_path_planer_b200/csrc/tma.cuh"
#include "../dqn-based-uav-3d_path_planer_b200/csrc/umma.cuh"
using namespace uavrl;

template <int N, int KP>
__global__ void __launch_bounds__(128) gemm_kernel(const float *A, const float *B, float *D, int K, int passes)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t SBO = umma_sbo(KP);
    unsigned char *Ahi = smem, *Alo = Ahi + umma_tile_bytes(128, KP);
    unsigned char *Bhi = Alo + umma_tile_bytes(128, KP), *Blo = Bhi + umma_tile_bytes(N, KP);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, 64);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    // operands -> canonical layout, hi/lo
    for (int i = tid; i < 128 * KP; i += 128) {
        const int r = i / KP, c = i % KP;
        float x = (c < K) ? A[r * K + c] : 0.f, hi, lo;
        tf32_split(x, hi, lo);
        *reinterpret_cast<float *>(Ahi + umma_off(r, c, SBO)) = hi;
        *reinterpret_cast<float *>(Alo + umma_off(r, c, SBO)) = lo;
    }
    for (int i = tid; i < N * KP; i += 128) {
        const int r = i / KP, c = i % KP;
        float x = (c < K) ? B[r * K + c] : 0.f, hi, lo;
        tf32_split(x, hi, lo);
        *reinterpret_cast<float *>(Bhi + umma_off(r, c, SBO)) = hi;
        *reinterpret_cast<float *>(Blo + umma_off(r, c, SBO)) = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(128, N);
        uint32_t acc = 0;
        for (int p = 0; p < passes; ++p) {
            const unsigned char *a = (p == 2) ? Alo : Ahi;
            const unsigned char *b = (p == 1) ? Blo : Bhi;
            for (int k = 0; k < KP / 8; ++k) {
                const uint64_t da = umma_desc(smem_u32(a) + k * 2 * kUmmaLBO, SBO);
                const uint64_t db = umma_desc(smem_u32(b) + k * 2 * kUmmaLBO, SBO);
                umma_tf32(tbase, da, db, idesc, acc);
                acc = 1;
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), N columns
    const int row = tid;
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 64);
}

int main()
{
    constexpr int N = 64, K = 100, KP = 104;
    std::vector<float> A(128 * K), B(N * K), D(128 * N);
    srand(1);
    for (auto &x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 4.f;
    for (auto &x : B) x = (rand() / (float)RAND_MAX - 0.5f) * 0.4f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = 2 * umma_tile_bytes(128, KP) + 2 * umma_tile_bytes(N, KP);
    cudaFuncSetAttribute(gemm_kernel<N, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int rc = 0;
    for (int passes = 1; passes <= 3; passes += 2) {
        cudaMemset(dD, 0, D.size() * 4);
        gemm_kernel<N, KP><<<1, 128, smem>>>(dA, dB, dD, K, passes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        double max_err = 0, max_ref = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * (double)B[n * K + k];
                max_err = fmax(max_err, fabs(s - D[r * N + n])); max_ref = fmax(max_ref, fabs(s));
            }
        printf("passes=%d  max_abs_err=%.3e  max_ref=%.3f  rel=%.3e  D[0][0]=%f D[5][7]=%f D[127][63]=%f\n", passes, max_err,
               max_ref, max_err / max_ref, D[0], D[5 * N + 7], D[127 * N + 63]);
        const double tol = (passes == 3) ? 2e-6 : 5e-3;
        if (!(max_err / max_ref < tol)) rc = 1;
    }
    printf(rc ? "FAIL\n" : "PASS\n");
    return rc;
}
