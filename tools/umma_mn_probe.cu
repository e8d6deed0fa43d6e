// Layout probe, round 2: MN-major tcgen05 operands for kind::tf32 with layout_type = 1 (SWIZZLE_128B_BASE32B), the only
// shared-memory layout CUTLASS accepts for MN-major 32-bit operands (cutlass/gemm/collective/builders/sm100_common.inl:92).
// Which SMEM float does the tensor core read for operand element (mn, k)?  SMEM is filled with values derived from the
// float index; the other operand is a K-major identity, so D = the probed values.  Two passes (index % 1024, index / 1024,
// both exact in TF32) give the index.  Expected (cute Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (32 mn x 4 k):(1, 32)):
//   byte(mn, k) = (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + ((((mn%32)/8) ^ (k%4))*32) + (mn%8)*4
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/mnprobe tools/umma_mn_probe.cu && /tmp/mnprobe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../dqn-based-uav-3d_path_planer_b200/csrc/tma.cuh"
#include "../dqn-based-uav-3d_path_planer_b200/csrc/umma.cuh"
using namespace uavrl;

constexpr int kFloats = 64 * 1024 / 4;

__global__ void __launch_bounds__(128) probe_kernel(float *D, int mode, int pass, uint32_t lbo, uint32_t sbo, uint32_t ltype)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *P = reinterpret_cast<float *>(smem);                       // probed operand region (1024-byte aligned)
    unsigned char *I = smem + kFloats * 4;                            // identity operand, K-major no swizzle, K_pad = 8
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, 32);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    for (int i = tid; i < kFloats; i += 128) P[i] = (float)(pass == 0 ? (i % 1024) : (i / 1024));
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
        dp |= (uint64_t)(ltype & 7u) << 61;
        const uint64_t di = umma_desc(smem_u32(I), ISBO);
        if (mode == 0) umma_tf32(tb, dp, di, umma_idesc_tf32(128, 16) | (1u << 15), 0u);      // A MN-major probed, B K-major identity
        else           umma_tf32(tb, di, dp, umma_idesc_tf32(128, 16) | (1u << 16), 0u);      // A K-major identity, B MN-major probed
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

static long expect(int mn, int k, uint32_t lbo, uint32_t sbo)
{
    return ((long)(mn / 32) * lbo + (long)(k / 4) * sbo + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4) / 4;
}

int main()
{
    float *dD;
    cudaMalloc(&dD, 128 * 16 * 4);
    const size_t smem = kFloats * 4 + 128 * 8 * 4 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // {lbo, sbo, layout_type}
    const uint32_t cfgs[][3] = { { 1024, 512, 1 }, { 512, 1024, 1 }, { 8192, 512, 1 }, { 1024, 512, 2 }, { 2048, 128, 0 } };
    for (int mode = 0; mode < 2; ++mode)
        for (auto &c : cfgs) {
            std::vector<float> lo(128 * 16), hi(128 * 16);
            for (int pass = 0; pass < 2; ++pass) {
                probe_kernel<<<1, 128, smem>>>(dD, mode, pass, c[0], c[1], c[2]);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d lbo %u sbo %u type %u: %s\n", mode, c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
                cudaMemcpy(pass ? hi.data() : lo.data(), dD, 128 * 16 * 4, cudaMemcpyDeviceToHost);
            }
            auto rd = [&](int mn, int k) { return mode == 0 ? (long)(hi[mn * 16 + k] * 1024 + lo[mn * 16 + k]) : (long)(hi[k * 16 + mn] * 1024 + lo[k * 16 + mn]); };
            const int nmn = mode == 0 ? 128 : 16;
            int ok = 0, tot = 0;
            for (int mn = 0; mn < nmn; ++mn) for (int k = 0; k < 8; ++k) { ++tot; ok += (rd(mn, k) == expect(mn, k, c[0], c[1])); }
            printf("mode %d (%s MN-major probed)  LBO %u  SBO %u  layout_type %u : %d / %d match the expected SW128_32B formula\n",
                   mode, mode ? "B" : "A", c[0], c[1], c[2], ok, tot);
            const int ms[] = { 0, 1, 2, 3, 4, 7, 8, 9, 15, 16, 24, 31, 32, 33, 40, 63, 64, 96, 127 };
            for (int m : ms) {
                if (m >= nmn) continue;
                printf("  mn=%3d:", m);
                for (int k = 0; k < 8; ++k) printf(" %6ld", rd(m, k));
                printf("\n");
            }
        }
    return 0;
}
