// umma.cuh -- tcgen05 (5th-gen tensor core) helpers for sm_100a, inline PTX: TMEM allocation, shared-memory
// matrix descriptors (K-major, no swizzle), kind::tf32 MMA issue, commit -> mbarrier, TMEM -> register loads.
//
// Operand layout used throughout (canonical K-major "interleave" layout, cute::UMMA::LayoutType::SWIZZLE_NONE):
//   core matrix = 8 rows x 16 bytes (4 fp32/tf32 values), stored as 128 contiguous bytes (row stride 16 B);
//   core matrices adjacent in K are LBO bytes apart, adjacent in M/N (next 8 rows) SBO bytes apart.
//   element (r, c) of a [rows][K] operand lives at   (r/8)*SBO + (c/4)*LBO + (r%8)*16 + (c%4)*4.
//   One kind::tf32 instruction consumes K = 8 (two 16-byte chunks); the next K step starts 2*LBO further.
#pragma once
#include <stdint.h>

namespace uavrl {

constexpr uint32_t kUmmaLBO = 128;                 // K-adjacent core matrices are contiguous

__host__ __device__ constexpr uint32_t umma_sbo(int k_pad) { return (uint32_t)(k_pad / 4) * 128u; }
__host__ __device__ constexpr uint32_t umma_tile_bytes(int rows, int k_pad) { return (uint32_t)(rows / 8) * umma_sbo(k_pad); }

// byte offset of element (r, c) inside an operand tile
__device__ __forceinline__ uint32_t umma_off(int r, int c, uint32_t sbo, uint32_t lbo = kUmmaLBO)
{
    return (uint32_t)(r >> 3) * sbo + (uint32_t)(c >> 2) * lbo + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 3) * 4u;
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 (no swizzle) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t sbo, uint32_t lbo = kUmmaLBO)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// ---- MN-major operands (the contraction index K is the SLOW one: element (mn, k) and (mn+1, k) are adjacent).
// kind::tf32 reads MN-major shared-memory operands only in layout_type 1 = SWIZZLE_128B_BASE32B (cute
// Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o (32 mn x 4 k):(1, 32); "for mn-major tf32 operands, SW128_32B is the only
// available smem layout", cutlass sm100_common.inl).  Measured on B200 with tools/umma_mn_probe.cu (every (mn, k) of an
// M=128 x K=8 A operand and an N=16 x K=8 B operand, three LBO/SBO settings):
//     byte(mn, k) = (mn/32)*LBO + (k/4)*SBO + (k%4)*128 + (((mn%32)/8) ^ (k%4))*32 + (mn%8)*4
// i.e. a [k][32 mn] row-major array of 128-byte rows whose four 32-byte chunks are XOR-swizzled with k%4; 4 k-rows form a
// 512-byte atom (1024-byte aligned base), SBO = distance between 4-k groups, LBO = distance between 32-wide mn blocks.
// A tensor stored [sample][feature] is therefore directly the MN-major operand of a product that contracts over samples.
constexpr uint32_t kMnAtom = 512;                  // bytes of one 4-k x 32-mn atom (the natural SBO when k groups are contiguous)
__device__ __forceinline__ uint32_t umma_mn_off(int mn, int k, uint32_t lbo, uint32_t sbo = kMnAtom)
{
    return (uint32_t)(mn >> 5) * lbo + (uint32_t)(k >> 2) * sbo + (uint32_t)(k & 3) * 128u +
           ((((uint32_t)(mn & 31) >> 3) ^ (uint32_t)(k & 3)) << 5) + (uint32_t)(mn & 7) * 4u;
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo = kMnAtom)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)1 << 61;                        // layout_type 1: SWIZZLE_128B_BASE32B
    return d;
}
constexpr uint32_t kUmmaAMajorMN = 1u << 15, kUmmaBMajorMN = 1u << 16;   // instruction-descriptor transpose bits

// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4) /* c = F32 */ | (2u << 7) /* a = TF32 */ | (2u << 10) /* b = TF32 */ |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)   // whole warp, ncols power of 2 >= 32
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)     // same warp that allocated
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T   (single thread issues)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes (rows), 32 consecutive fp32 columns starting at taddr
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Accumulator read for the concatenated 3xTF32 scheme: columns [c, c+32) plus, when second_off != 0, the partner block
// [second_off + c, ...) (the hi*lo product); both loads are in flight before the single wait.
__device__ __forceinline__ void tmem_ld32_sum(uint32_t taddr, uint32_t second_off, float (&v)[32])
{
    if (second_off == 0) { tmem_ld32(taddr, v); return; }
    uint32_t r[32], s[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]),
                   "=r"(s[8]), "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15]),
                   "=r"(s[16]), "=r"(s[17]), "=r"(s[18]), "=r"(s[19]), "=r"(s[20]), "=r"(s[21]), "=r"(s[22]), "=r"(s[23]),
                   "=r"(s[24]), "=r"(s[25]), "=r"(s[26]), "=r"(s[27]), "=r"(s[28]), "=r"(s[29]), "=r"(s[30]), "=r"(s[31])
                 : "r"(taddr + second_off) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + __uint_as_float(s[i]);
}

// 3xTF32 split: hi = x rounded to TF32 (10-bit mantissa), lo = x - hi (exact).  hi*hi + hi*lo + lo*hi
// reproduces the fp32 product to ~2^-21.
// Round-to-nearest, ties away (cvt.rna.tf32.f32) on the bit pattern: add half an ulp of the 10-bit mantissa to the magnitude
// and clear the 13 dropped bits -- two integer instructions.  ptxas expands the cvt into the same two plus an |x| < inf
// test and a select per element (inf/NaN kept verbatim); for every finite x that does not round up to inf the result is
// bit-identical, inf stays inf, and the epilogues run this 32 times per thread per layer.
__device__ __forceinline__ void tf32_split(float x, float &hi, float &lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}

// Issue the three TF32 products of one layer (single thread), K = 8 * ksteps.
//   concat: the B operand's hi and lo blocks are contiguous in SMEM ([B_hi ; B_lo] = 2N rows), so
//     D[:, 0:2N]  = A_hi * [B_hi ; B_lo]^T   (one N = 2N instruction per K step: A_hi is read from SMEM once)
//     D[:, 0:N]  += A_lo * B_hi^T
//   -> 2 instructions per K step instead of 3; the epilogue adds D[:, c] + D[:, N + c] (tmem_ld32_sum).
//   !concat: hi*hi + hi*lo + lo*hi accumulate into the same N columns (3 instructions per K step).
__device__ __forceinline__ void issue_3xtf32(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, int M, int N,
                                             int ksteps, bool concat, uint32_t lbo = kUmmaLBO)
{
    const uint64_t kStep = (uint64_t)((2 * lbo) >> 4);
    const uint32_t idesc = umma_idesc_tf32(M, N);
    uint64_t da = a_hi, db = b_hi;
    if (concat) {
        const uint32_t wide = umma_idesc_tf32(M, 2 * N);
        umma_tf32(d, da, db, wide, 0u);
        for (int k = 1; k < ksteps; ++k) { da += kStep; db += kStep; umma_tf32(d, da, db, wide, 1u); }
    } else {
        umma_tf32(d, da, db, idesc, 0u);
        for (int k = 1; k < ksteps; ++k) { da += kStep; db += kStep; umma_tf32(d, da, db, idesc, 1u); }
        da = a_hi; db = b_lo;
        for (int k = 0; k < ksteps; ++k) { umma_tf32(d, da, db, idesc, 1u); da += kStep; db += kStep; }
    }
    da = a_lo; db = b_hi;
    for (int k = 0; k < ksteps; ++k) { umma_tf32(d, da, db, idesc, 1u); da += kStep; db += kStep; }
}

// Stacked 3xTF32 for tiles of at most 64 real rows R (the M = 128 instruction costs the same however many rows are real, so the
// idle rows carry the second operand half): the A operand holds A_hi in rows [0, R) and A_lo in rows [R, 2R) of ONE operand,
// B = [B_hi ; B_lo] (2N rows, contiguous): a single instruction per K step yields
//     rows [0, R):   columns [0, N) = hi*hi,   columns [N, 2N) = hi*lo
//     rows [R, 2R):  columns [0, N) = lo*hi    (columns [N, 2N) = lo*lo: not needed)
// -- half the instructions of issue_3xtf32.  The epilogue adds D[r][c] + D[r][N + c] + D[R + r][c]; the third term sits in
// another TMEM lane quadrant and travels through a shared-memory scratch (stack_park_lo / stack_add_lo).
__device__ __forceinline__ void issue_3xtf32_stacked(uint32_t d, uint64_t a, uint64_t b, int M, int N, int ksteps, uint32_t lbo = kUmmaLBO)
{
    const uint64_t kStep = (uint64_t)((2 * lbo) >> 4);
    const uint32_t wide = umma_idesc_tf32(M, 2 * N);
    umma_tf32(d, a, b, wide, 0u);
    for (int k = 1; k < ksteps; ++k) { a += kStep; b += kStep; umma_tf32(d, a, b, wide, 1u); }
}
constexpr int kLoLd = 68;                          // floats per scratch row: 64 + 4, so 16-byte accesses of 8 lanes hit 32 banks
// a warp of the lo quadrant(s): park D[R + r][c0 .. c0 + 32) (r = its row within the lo block) in the scratch
__device__ __forceinline__ void stack_park_lo(uint32_t taddr, int c0, float *s_lo, int r)
{
    float v[32];
    tmem_ld32(taddr + (uint32_t)c0, v);
    float4 *dst = reinterpret_cast<float4 *>(s_lo + r * kLoLd + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void stack_add_lo(float (&v)[32], const float *s_lo, int r, int c0)
{
    const float4 *src = reinterpret_cast<const float4 *>(s_lo + r * kLoLd + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = src[j];
        v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
}

// The three TF32 products for operands that are BOTH MN-major (layout above), contraction over K = 8 * ksteps.  The k-th
// step starts 2 * sbo further (two 4-k groups).  concat: [B_hi ; B_lo] are adjacent mn blocks (lbo apart), see issue_3xtf32.
// accumulate != 0: add to what the accumulator already holds (a CTA summing several sample chunks into one partial).
__device__ __forceinline__ void issue_3xtf32_mn(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, int M, int N,
                                                int ksteps, bool concat, uint32_t accumulate = 0u, uint32_t sbo = kMnAtom)
{
    const uint64_t kStep = (uint64_t)((2 * sbo) >> 4);
    const uint32_t mn = kUmmaAMajorMN | kUmmaBMajorMN;
    const uint32_t idesc = umma_idesc_tf32(M, N) | mn;
    uint64_t da = a_hi, db = b_hi;
    if (concat) {
        const uint32_t wide = umma_idesc_tf32(M, 2 * N) | mn;
        umma_tf32(d, da, db, wide, accumulate);
        for (int k = 1; k < ksteps; ++k) { da += kStep; db += kStep; umma_tf32(d, da, db, wide, 1u); }
    } else {
        umma_tf32(d, da, db, idesc, accumulate);
        for (int k = 1; k < ksteps; ++k) { da += kStep; db += kStep; umma_tf32(d, da, db, idesc, 1u); }
        da = a_hi; db = b_lo;
        for (int k = 0; k < ksteps; ++k) { umma_tf32(d, da, db, idesc, 1u); da += kStep; db += kStep; }
    }
    da = a_lo; db = b_hi;
    for (int k = 0; k < ksteps; ++k) { umma_tf32(d, da, db, idesc, 1u); da += kStep; db += kStep; }
}

}  // namespace uavrl
