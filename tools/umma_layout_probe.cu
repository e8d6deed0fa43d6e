// Layout probe for MN-major tcgen05 operands (kind::tf32, no swizzle): which SMEM float does the tensor core read for
// A element (m, k)?  SMEM is filled with small integers derived from the float index; B is a K-major identity, so
// D[m][k] = A(m, k) = the probed value.  Two passes (index % 1024 and index / 1024, both exact in TF32) give the index.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/probe tools/umma_layout_probe.cu && /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../dqn-based-uav-3d_path_planer_b200/csrc/tma.cuh"
#include "../dqn-based-uav-3d_path_planer_b200/csrc/umma.cuh"
using namespace uavrl;
// RESULT on B200 (gpurun_out/probe.txt, round 1): the K-major sanity mode reads element (mn, k) at float index
// mn*4 + (k%4) + (k/4)*32 (LBO 128 B), as umma.cuh assumes; with a_major / b_major = 1 and layout_type 0 (no swizzle) the
// instruction returns ZEROS for every (mn, k) and both LBO/SBO assignments: kind::tf32 does not read an unswizzled
// MN-major operand.  The weight-gradient kernel therefore transposes in registers (tc_train.cu dw_build_transposed).

constexpr int kFloats = 72 * 1024 / 4;          // probed region: 72 KB (M = 128 reads 32 groups x 2 KB = 64 KB)

// mode 0: A MN-major probed (B K-major identity);  mode 1: B MN-major probed (A K-major identity rows 0..7)
__global__ void __launch_bounds__(128) probe_kernel(float *D, int mode, int pass, uint32_t lbo, uint32_t sbo)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *P = reinterpret_cast<float *>(smem);                       // probed operand region
    unsigned char *I = smem + kFloats * 4;                            // identity operand, K-major, K_pad = 8: [rows][8]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, 32);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = tid; i < kFloats; i += 128) P[i] = (float)(pass == 0 ? (i % 1024) : (i / 1024));
    if (mode == 2) {                                                  // sanity: K-major probe, K_pad = 8 (known-good path)
        // nothing else: descriptor below uses LBO 128 / SBO 256 with a_major = 0
    }
    constexpr uint32_t ISBO = umma_sbo(8);
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i / 8, c = i % 8;
        *reinterpret_cast<float *>(I + umma_off(r, c, ISBO)) = (r == c) ? 1.f : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        uint64_t dp = 0;
        dp |= (uint64_t)((smem_u32(P) >> 4) & 0x3FFFu);
        dp |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
        dp |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
        dp |= (uint64_t)1 << 46;
        const uint64_t di = umma_desc(smem_u32(I), ISBO);
        if (mode == 2) umma_tf32(tb, dp, di, umma_idesc_tf32(128, 16), 0u);                   // A K-major (sanity)
        else if (mode == 0) umma_tf32(tb, dp, di, umma_idesc_tf32(128, 16) | (1u << 15), 0u);      // A MN-major, B K-major
        else           umma_tf32(tb, di, dp, umma_idesc_tf32(128, 16) | (1u << 16), 0u);      // A K-major identity, B MN-major
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    float v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), v);
    for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * 16 + j] = v[j];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 32);
}

int main()
{
    float *dD;
    cudaMalloc(&dD, 128 * 16 * 4);
    const size_t smem = kFloats * 4 + 128 * 8 * 4 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const uint32_t cfgs[][2] = { { 128, 2048 }, { 2048, 128 }, { 128, 256 } };
    const int modes[] = { 2, 0, 1 };
    for (int mode : modes)
        for (auto &c : cfgs) {
            if ((mode == 2) != (c[1] == 256)) continue;
            std::vector<float> lo(128 * 16), hi(128 * 16);
            for (int pass = 0; pass < 2; ++pass) {
                probe_kernel<<<1, 128, smem>>>(dD, mode, pass, c[0], c[1]);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d lbo %u sbo %u: %s\n", mode, c[0], c[1], cudaGetErrorString(e)); return 1; }
                cudaMemcpy(pass ? hi.data() : lo.data(), dD, 128 * 16 * 4, cudaMemcpyDeviceToHost);
            }
            printf("mode %d (%s probed; 2 = K-major sanity)  LBO field %u  SBO field %u : float index read for (mn, k)\n", mode, mode ? "B" : "A", c[0], c[1]);
            if (mode != 1) {
                const int ms[] = { 0, 1, 2, 3, 4, 5, 8, 16, 32, 64, 100, 127 };
                for (int m : ms) {
                    printf("  mn=%3d:", m);
                    for (int k = 0; k < 8; ++k) printf(" %6d", (int)(hi[m * 16 + k] * 1024 + lo[m * 16 + k]));
                    printf("\n");
                }
            } else {
                // D[m][n] = sum_k I[m][k] B[n][k] = B(n, k=m) for m < 8
                for (int n = 0; n < 16; ++n) {
                    printf("  mn=%3d:", n);
                    for (int k = 0; k < 8; ++k) printf(" %6d", (int)(hi[k * 16 + n] * 1024 + lo[k * 16 + n]));
                    printf("\n");
                }
            }
        }
    return 0;
}
