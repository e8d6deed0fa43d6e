// tc_train.cu -- the TD update's forward + backward on the tensor cores (tcgen05 / TMEM, 3xTF32), sm_100a.
//
// Replaces the arithmetic of Trainer.update (Trainer/DuelingDQN_Trainer.py:164-180, DDQN_Trainer.py:93-107,
// DQN_Trainer.py:107-126): Q(s).gather(a), MSE against the TD target, loss.backward().
//
//   tc_train_kernel  one CTA = R (32/64) sampled transitions.  Forward chain exactly like tc_forward.cu, then the
//                    loss / dLoss/dQ in the head epilogue, then the dX chain with the TRANSPOSED weight blocks of the
//                    training image:  dH_l = dZ_{l+1} * W_{l+1}  (A = dZ rows, B = W^T K-major), ReLU mask applied
//                    in the epilogue.  Hidden activations and every dZ are written once to a per-sample scratch
//                    (L2 resident: 4096 x ~0.9 KB) for the weight-gradient pass.
//   tc_dw_kernel     split-K weight gradients: CTA (layer l, chunk of 128 samples) computes
//                    dW_l^T [in+1 x out] = [act_l ; 1]^T (in+1 x 128) * dZ_l (128 x out)  -- the extra all-ones
//                    row yields the bias gradient for free -- and stores its slice of partial `chunk`.
//                    The contraction runs over SAMPLES, and both operands are stored [sample][feature]: exactly the
//                    MN-major tcgen05 operand layout (umma.cuh), so the rows go from global memory to SMEM with plain
//                    16-byte copies -- no transposition.  Small batches (grid <= SMs): the same kernel then meets at a
//                    grid barrier and performs the partial reduction + Adam + image refresh itself (one launch less);
//                    otherwise reduce_adam_kernel sums the B/128 partials.
// All products use the hi*hi + hi*lo + lo*hi TF32 split (fp32-grade, see tc_forward.cu).
#include <string.h>
#include <stdlib.h>

#include <atomic>

#include "tc_forward.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace uavrl {

constexpr int kDwChunk = 128;          // samples reduced by one dW CTA (the MMA's K extent)

struct TcTrainArgs {
    const unsigned char *img;          // local network, training image (forward blocks + biases + transposed blocks)
    BatchSrc src;
    int32_t B, R, n_tiles;
    const float *y;                    // [B] TD targets
    float inv_global_b;
    float *act_buf, *dz_buf;           // per-sample scratch rows
    float *loss_partials;              // [grid]
    int32_t loss_kind;                 // 0 MSE (reference), 1 Huber (delta = 1)
    // fused TD target (fused_td = 1; only when every CTA owns exactly one tile): the same CTA first evaluates the target
    // network (double DQN: the local network for a*, then the target network) on the tile's NEXT states and keeps
    // y = r + gamma * next_q * (1 - d) in shared memory -- one launch instead of two or three, no y round trip
    int32_t fused_td, algo;
    float gamma;
    const unsigned char *img_target;   // target network, forward image
    long long *trace;                  // debug (UAVRL_TC_TRACE): CTA 0 / thread 0 stage timestamps
};
#define TR_TRACE(slot) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[slot] = clock64(); } while (0)

struct TcDwArgs {
    BatchSrc src;
    int32_t B, n_chunks, P;
    int32_t n_slices;                  // CTAs per layer; slice i accumulates chunks i, i + n_slices, ... into partial i
    const float *act_buf, *dz_buf;
    float *partials;                   // [n_slices][P]
    long long *trace;                  // debug (UAVRL_TC_TRACE): CTA 0 / thread 0 stage timestamps
    // fused optimiser tail (fuse_adam = 1, only when every CTA of the grid is resident at once): the partials travel as 8-byte
    // words {epoch : value} (part64 [n_slices][P]); a reader polls the word itself -- no grid barrier, no fence, no flag
    int32_t fuse_adam;
    unsigned long long *part64;
    uint32_t ll_epoch;                 // this launch's tag (never 0: the buffer starts zeroed)
    AdamArgs adam;
    AdamPtrs ptrs;
};
#define DW_TRACE(slot) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[slot] = clock64(); } while (0)

// Gather R rows (pointers in rows[]) into the layer-0 A operand.  All of a thread's loads are issued before any
// is consumed (4 in flight), so the gather costs one L2 round trip instead of one per chunk.  Item i = (chunk j = i / R,
// row r = i % R; R is a power of two): consecutive lanes take consecutive rows of the same 16-byte chunk, so a quarter-warp's
// 16-byte stores cover one whole core matrix column = 128 contiguous bytes (lanes walking along a row would all hit the
// same 4 banks) and the index needs no division.
__device__ __forceinline__ void build_a0(const float *const *rows, int R, int in_dim, int K0, unsigned char *Ahi, unsigned char *Alo, bool stack)
{
    const int chunks = K0 / 4, total = R * chunks, lgR = 31 - __clz(R);
    const uint32_t sbo = umma_sbo(K0);
    for (int i0 = threadIdx.x; i0 < total; i0 += 4 * kTcThreads) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kTcThreads;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total) {
                const int r = i & (R - 1), j = i >> lgR;
                if (rows[r] && 4 * j < in_dim) v[u] = __ldg(reinterpret_cast<const float4 *>(rows[r]) + j);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kTcThreads;
            if (i < total) {
                const int r = i & (R - 1), j = i >> lgR;
                float4 h, l;
                tf32_split(v[u].x, h.x, l.x); tf32_split(v[u].y, h.y, l.y); tf32_split(v[u].z, h.z, l.z); tf32_split(v[u].w, h.w, l.w);
                const uint32_t off = umma_off(r, 4 * j, sbo);
                *reinterpret_cast<float4 *>(Ahi + off) = h;
                if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + r, 4 * j, sbo)) = l;   // stacked 3xTF32: A_lo in rows [R, 2R)
                else *reinterpret_cast<float4 *>(Alo + off) = l;
            }
        }
    }
}

// The same gather in two halves for a tile of at most 4 * kTcThreads items (R = 32: 832): the loads are issued long before the
// operand buffer is free (the fused TD pre-pass runs in between), so the HBM latency of the sampled rows is off the chain.
__device__ __forceinline__ void a0_load(const float *const *rows, int R, int in_dim, int K0, float4 (&v)[4])
{
    const int total = R * (K0 / 4), lgR = 31 - __clz(R);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * kTcThreads;
        const int r = i & (R - 1), j = i >> lgR;
        const float *rp = (i < total) ? rows[r] : nullptr;
        v[u] = (rp && 4 * j < in_dim) ? __ldg(reinterpret_cast<const float4 *>(rp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// want_fresh = false: every row NOT flagged in fresh[] (everything else zero) -- may run before the dependent-launch wait;
// want_fresh = true: only the flagged rows, read from L2 (the env step of this iteration has just written them), into the
// same registers.
__device__ __forceinline__ void a0_load_sel(const float *const *rows, const uint8_t *fresh, bool want_fresh, int R, int in_dim, int K0, float4 (&v)[4])
{
    const int total = R * (K0 / 4), lgR = 31 - __clz(R);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * kTcThreads;
        const int r = i & (R - 1), j = i >> lgR;
        const float *rp = (i < total) ? rows[r] : nullptr;
        const bool take = rp && 4 * j < in_dim && ((fresh[r] != 0) == want_fresh);
        if (!want_fresh) v[u] = take ? __ldg(reinterpret_cast<const float4 *>(rp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        else if (take) v[u] = __ldcg(reinterpret_cast<const float4 *>(rp) + j);
    }
}
__device__ __forceinline__ void a0_store(const float4 (&v)[4], int R, int K0, unsigned char *Ahi, unsigned char *Alo, bool stack)
{
    const int total = R * (K0 / 4), lgR = 31 - __clz(R);
    const uint32_t sbo = umma_sbo(K0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = threadIdx.x + u * kTcThreads;
        if (i < total) {
            const int r = i & (R - 1), j = i >> lgR;
            float4 h, l;
            tf32_split(v[u].x, h.x, l.x); tf32_split(v[u].y, h.y, l.y); tf32_split(v[u].z, h.z, l.z); tf32_split(v[u].w, h.w, l.w);
            const uint32_t off = umma_off(r, 4 * j, sbo);
            *reinterpret_cast<float4 *>(Ahi + off) = h;
            if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + r, 4 * j, sbo)) = l;
            else *reinterpret_cast<float4 *>(Alo + off) = l;
        }
    }
}

__device__ __forceinline__ uint32_t nl_split(const TcNet &tc) { return tc.n_layers > 1 ? (uint32_t)tc.L[1].hi_off : (uint32_t)tc.img_bytes; }

// STACK / NPRE / DUELING are compile-time so that the kernel a configuration runs carries no code of the others: a third of the
// live warps' stall samples of the generic kernel were instruction-fetch stalls (143 KB of code, executed once per CTA).
// NPRE = forward-only TD passes ahead of the training chain (0: y comes from stand-alone passes, 1: DQN, 2: double DQN).
template <bool STACK, int NPRE, bool DUELING>
__global__ void __launch_bounds__(kTcThreads, 1) tc_train_kernel(TcNet tc, TcTrainArgs a)
{
    TR_TRACE(0);
    extern __shared__ __align__(1024) unsigned char smem[];
    const int R = a.R;
    const uint32_t a_bytes = (uint32_t)(R / 8) * umma_sbo(tc.max_k);
    unsigned char *Ahi = smem, *Alo = smem + a_bytes, *W = smem + 2 * a_bytes;
    // stacked 3xTF32 (umma.cuh; R is 32 or 64 here): A_lo occupies rows [R, 2R) of the operand that starts at Ahi (which runs on
    // into the Alo buffer); the lo*hi accumulator block reaches the epilogue warps through a scratch behind the weight image
    constexpr bool stack = STACK;                                    // = tc.concat && tc.dstride <= 128 (launch_tc_train)
    float *s_lo = reinterpret_cast<float *>(W + tc.train_img_bytes);
    __shared__ uint64_t wbar, wbar2, wbar3, mbar;           // fused TD: training image in three pieces (below)
    __shared__ uint32_t tmem_base_s;
    __shared__ const float *rows[kTcTile];
    __shared__ const float *rows2[kTcTile];                  // fused TD: next-state rows
    __shared__ int s_act[kTcTile], s_astar[kTcTile];
    __shared__ uint8_t s_fresh[kTcTile];                     // fused TD: the next-state row is being written by this iteration's env step
    __shared__ float s_y[kTcTile], s_rew[kTcTile], s_done[kTcTile];
    __shared__ float s_loss;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
    // control warp (the last one): barrier init, weight copies, TMEM allocation -- concurrently with warp 0 resolving the tile's
    // samples; the first CTA-wide barrier (behind the sample table) publishes all of it (tc_forward.cu)
    constexpr int kCtl = kTcThreads - 32;
    if (tid == 0) s_loss = 0.f;
    constexpr bool fused = NPRE > 0;
    constexpr int n_pre = NPRE;                                      // forward-only passes ahead of the training chain
    // PDL.  Unfused: the training image was written by the previous optimiser kernel (>= 2 kernels back: a TD pass always
    // precedes this kernel), so it is fetched before the wait, and so are the first tile's sampled rows and actions
    // (replay frames / actions were written by the env step and the act kernel, also >= 2 back); only y is the
    // predecessor's output.  Fused: the predecessor is the env step, which writes the newest frame's rows, rewards and
    // flags -- only the first pass's weight image (optimiser kernel, >= 2 back) is fetched before the wait.
    uint32_t wphase = 0;
    if (tid == kCtl) {
        mbar_init(&wbar, 1); mbar_init(&wbar2, 1); mbar_init(&wbar3, 1); mbar_init(&mbar, 1); fence_barrier_init();
        fence_proxy_async();
        if (!fused) bulk_g2s_chunked(W, a.img, (uint32_t)tc.train_img_bytes, &wbar);
        else {
            bulk_g2s_chunked(W, (n_pre == 2) ? a.img : a.img_target, (uint32_t)tc.img_bytes, &wbar);
            // the transposed blocks of the training image (dX chain) lie behind the forward image the TD passes use: they are
            // fetched now (written by the optimiser step, >= 2 kernels back) and first waited for in front of the dX chain
            if (tc.train_img_bytes > tc.img_bytes)
                bulk_g2s_chunked(W + tc.img_bytes, a.img + tc.img_bytes, (uint32_t)(tc.train_img_bytes - tc.img_bytes), &wbar3);
        }
    }
    // fused: after the TD passes the forward part of the training image comes in two pieces -- layer 0's block on wbar (all the
    // first MMA needs), the other layers + biases on wbar2 (first waited for in layer 0's epilogue)
    const uint32_t w_split = nl_split(tc);
    if (warp == kCtl / 32) { __syncwarp(); tmem_alloc(&tmem_base_s, (uint32_t)tc.tmem_cols); tc_fence_before(); }
    uint32_t tmem = 0;
    TR_TRACE(1);
    bool waited = false;
    const float *bias_all = reinterpret_cast<const float *>(W + tc.bias_base);

    uint32_t pkey[4];
    Philox::gen(a.src.key, a.src.epoch, 0x5A17ull, pkey);
    uint32_t mphase = 0;
    bool wready = false;
    const int nl = tc.n_layers;
    const int row = quad * 32 + lane;
    const bool live = quad * 32 < R;
    // ReLU' for the dX chain: bit j of hmK = (H_K[row][half * 32 + j] > 0), kept from the forward epilogue of the same thread
    // (hidden layers are at most 64 wide here: one 32-column chunk per thread) -- no reload of H from global memory
    uint32_t hm1 = 0u, hm2 = 0u, hm3 = 0u, hm4 = 0u;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int base = tile * R;
        // where the tile's rows are: index arithmetic only (resolve_rows), so with the fused TD pass it runs BEFORE the wait for
        // the env step; the transitions' action / reward / done are requested right after the wait and parked in shared memory
        // once the first gather has been issued (they are first read in a head epilogue)
        int64_t my_slot = -1;
        if (tid < R) {
            const int b = base + tid;
            const float *p = nullptr, *p2 = nullptr;
            bool fr = false;
            if (b < a.B) my_slot = resolve_rows(a.src, b, tc.in_dim, pkey, p, p2, &fr);
            rows[tid] = p;
            if (fused) { rows2[tid] = p2; s_astar[tid] = 0; s_fresh[tid] = fr ? 1 : 0; }
        }
        tc_fence_before();
        __syncthreads();                                         // the sample table -- and, first time round, barriers + TMEM base
        tc_fence_after();
        tmem = tmem_base_s;
        // Fused TD (one tile per CTA, at most 4 items per thread): BOTH gathers are requested before the wait for the env step --
        // the training rows (never younger than the previous iteration) and every next-state row outside the frame that env step
        // is writing (all but ~N/count of them); the few fresh ones follow after the wait from L2.  The HBM latency of the
        // sampled rows is then hidden behind the predecessor instead of heading this kernel's chain.
        const bool early_rows = n_pre > 0 && R * (tc.L[0].K_pad / 4) <= 4 * kTcThreads;
        float4 vmain[4], vnext[4];
        if (early_rows) {
            a0_load_sel(rows2, s_fresh, false, R, tc.in_dim, tc.L[0].K_pad, vnext);
            a0_load(rows, R, tc.in_dim, tc.L[0].K_pad, vmain);
        }
        if (fused && !waited) { pdl_wait(); pdl_trigger(); waited = true; }
        int m_act = 0; float m_rew = 0.f, m_done = 0.f;
        if (my_slot >= 0) load_meta(a.src, my_slot, m_act, m_rew, m_done);
        if (early_rows) a0_load_sel(rows2, s_fresh, true, R, tc.in_dim, tc.L[0].K_pad, vnext);
        auto park_meta = [&]() { if (tid < R) { s_act[tid] = m_act; if (fused) { s_rew[tid] = m_rew; s_done[tid] = m_done; } } };
        TR_TRACE(2);
        // ---------------- fused TD target: forward-only pass(es) on the next states (tc_forward.cu's chain and head)
        for (int pass = 0; pass < n_pre; ++pass) {
            if (pass > 0 && tid == kTcThreads - 32) {                          // the target image replaces the local one (all its readers are done)
                fence_proxy_async();
                bulk_g2s_chunked(W, a.img_target, (uint32_t)tc.img_bytes, &wbar);
            }
            if (pass == 0 && early_rows) a0_store(vnext, R, tc.L[0].K_pad, Ahi, Alo, stack);
            else build_a0(rows2, R, tc.in_dim, tc.L[0].K_pad, Ahi, Alo, stack);
            if (pass == 0) { park_meta(); TR_TRACE(3); }
            mbar_wait(&wbar, wphase); wphase ^= 1;
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            if (pass == 0) TR_TRACE(4);
            const bool td_pass = (pass == n_pre - 1);
            for (int l = 0; l < nl; ++l) {
                const TcLayer T = tc.L[l];
                const uint32_t sbo = umma_sbo(T.K_pad);
                const uint32_t dcol = (uint32_t)(l & 1) * (uint32_t)tc.dstride;
                const uint32_t second = tc.concat ? (uint32_t)T.N_pad : 0u;
                if (tid == 0) {
                    if (stack) issue_3xtf32_stacked(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(W + T.hi_off), sbo), kTcTile, T.N_pad, T.K_pad / 8);
                    else issue_3xtf32(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(Alo), sbo),
                                      umma_desc(smem_u32(W + T.hi_off), sbo), umma_desc(smem_u32(W + T.lo_off), sbo), kTcTile, T.N_pad,
                                      T.K_pad / 8, tc.concat != 0);
                    umma_commit(&mbar);
                }
                mbar_wait(&mbar, mphase);
                mphase ^= 1;
                tc_fence_after();
                const float *bias = bias_all + T.bias_off;
                const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + dcol;
                float vpre[32];                                      // this warp's accumulator chunk, loaded while the lo warps park theirs
                if (stack) {                                       // lo*hi block (rows [R, 2R)) -> scratch
                    if (quad * 32 >= R && quad * 32 < 2 * R)
                        for (int c0 = half * 32; c0 < T.N_pad; c0 += 64) stack_park_lo(taddr, c0, s_lo, row - R);
                    if (live && half * 32 < T.N_pad) tmem_ld32_sum(taddr + (uint32_t)(half * 32), second, vpre);
                    __syncthreads();
                }
                if (l + 1 < nl) {
                    const uint32_t sbon = umma_sbo(T.N_pad);
                    for (int c0 = half * 32; live && c0 < T.N_pad; c0 += 64) {
                        float v[32];
                        if (stack) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = vpre[j];
                            stack_add_lo(v, s_lo, row, c0);
                        } else tmem_ld32_sum(taddr + (uint32_t)c0, second, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float4 x, h, lo4;
                            x.x = fmaxf(v[4 * j + 0] + bias[c0 + 4 * j + 0], 0.f); x.y = fmaxf(v[4 * j + 1] + bias[c0 + 4 * j + 1], 0.f);
                            x.z = fmaxf(v[4 * j + 2] + bias[c0 + 4 * j + 2], 0.f); x.w = fmaxf(v[4 * j + 3] + bias[c0 + 4 * j + 3], 0.f);
                            tf32_split(x.x, h.x, lo4.x); tf32_split(x.y, h.y, lo4.y); tf32_split(x.z, h.z, lo4.z); tf32_split(x.w, h.w, lo4.w);
                            const uint32_t off = umma_off(row, c0 + 4 * j, sbon);
                            *reinterpret_cast<float4 *>(Ahi + off) = h;
                            if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + row, c0 + 4 * j, sbon)) = lo4;
                            else *reinterpret_cast<float4 *>(Alo + off) = lo4;
                        }
                    }
                } else if (half == 0 && live) {
                    float q[32];
                    if (stack) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] = vpre[j];
                        stack_add_lo(q, s_lo, row, 0);
                    } else tmem_ld32_sum(taddr, second, q);
                    const int nA = tc.n_actions;
#pragma unroll
                    for (int j = 0; j < 32; ++j) q[j] += bias[j];
                    if (DUELING) {                                 // Q = V + A - mean(A)  (BaseCNN.py:138)
                        float sA = 0.f, V = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < nA) sA += q[j]; if (j == nA) V = q[j]; }
                        const float mean = sA / (float)nA;
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] = V + q[j] - mean;
                    }
                    int best = 0; float bv = q[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) if (j < nA && q[j] > bv) { bv = q[j]; best = j; }
                    if (!td_pass) s_astar[row] = best;                // DDQN_Trainer.py:94
                    else {
                        float nq = bv;                                // DQN_Trainer.py:109
                        if (n_pre == 2) {                             // DDQN_Trainer.py:95: gather at a*
                            const int as = s_astar[row];
                            nq = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (j == as) nq = q[j];
                        }
                        s_y[row] = s_rew[row] + (a.gamma * nq * (1.f - s_done[row]));      // :99 / :114 / :171
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                tc_fence_after();
                if (pass == 0) TR_TRACE(5 + l);
            }
        }
        if (fused && tid == kTcThreads - 32) {                                  // the training image (every reader of the TD image is done)
            fence_proxy_async();
            bulk_g2s_chunked(W, a.img, w_split, &wbar);
            if (w_split < (uint32_t)tc.img_bytes) bulk_g2s_chunked(W + w_split, a.img + w_split, (uint32_t)tc.img_bytes - w_split, &wbar2);
        }
        if (early_rows) a0_store(vmain, R, tc.L[0].K_pad, Ahi, Alo, stack);
        else build_a0(rows, R, tc.in_dim, tc.L[0].K_pad, Ahi, Alo, stack);
        TR_TRACE(9);
        if (!waited) { pdl_wait(); pdl_trigger(); waited = true; }
        if (n_pre == 0) park_meta();
        if (!fused && tid < R) s_y[tid] = (base + tid < a.B) ? a.y[base + tid] : 0.f;      // visible after the barrier below
        if (fused) { mbar_wait(&wbar, wphase); wphase ^= 1; }
        else if (!wready) { mbar_wait(&wbar, 0); wready = true; }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        TR_TRACE(10);
        const int gb = base + row;                          // this thread's sample (valid when live && gb < B)
        const bool mine = live && gb < a.B;

        // ---------------- forward chain
        for (int l = 0; l < nl; ++l) {
            const TcLayer T = tc.L[l];
            const uint32_t sbo = umma_sbo(T.K_pad);
            const uint32_t dcol = (uint32_t)(l & 1) * (uint32_t)tc.dstride;
            const uint32_t second = tc.concat ? (uint32_t)T.N_pad : 0u;
            if (tid == 0) {
                if (stack) issue_3xtf32_stacked(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(W + T.hi_off), sbo), kTcTile, T.N_pad, T.K_pad / 8);
                else issue_3xtf32(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(Alo), sbo),
                                  umma_desc(smem_u32(W + T.hi_off), sbo), umma_desc(smem_u32(W + T.lo_off), sbo), kTcTile, T.N_pad,
                                  T.K_pad / 8, tc.concat != 0);
                umma_commit(&mbar);
            }
            mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc_fence_after();
            if (fused && l == 0 && w_split < (uint32_t)tc.img_bytes) mbar_wait(&wbar2, 0);     // biases + the later layers' weights
            if (l < 4) TR_TRACE(27 + l);
            const float *bias = bias_all + T.bias_off;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + dcol;
            float vpre[32];                                      // this warp's accumulator chunk, loaded while the lo warps park theirs
            if (stack) {                                           // lo*hi block (rows [R, 2R)) -> scratch
                if (quad * 32 >= R && quad * 32 < 2 * R)
                    for (int c0 = half * 32; c0 < T.N_pad; c0 += 64) stack_park_lo(taddr, c0, s_lo, row - R);
                if (live && half * 32 < T.N_pad) tmem_ld32_sum(taddr + (uint32_t)(half * 32), second, vpre);
                __syncthreads();
            }
            if (l + 1 < nl) {
                const uint32_t sbon = umma_sbo(T.N_pad);
                float *act_row = a.act_buf + (size_t)gb * tc.act_stride + tc.L[l + 1].act_off;
                for (int c0 = half * 32; live && c0 < T.N_pad; c0 += 64) {
                    float v[32];
                    if (stack) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = vpre[j];
                        stack_add_lo(v, s_lo, row, c0);
                    } else tmem_ld32_sum(taddr + (uint32_t)c0, second, v);
                    uint32_t mk = 0u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 x, h, lo4;
                        x.x = fmaxf(v[4 * j + 0] + bias[c0 + 4 * j + 0], 0.f); x.y = fmaxf(v[4 * j + 1] + bias[c0 + 4 * j + 1], 0.f);
                        x.z = fmaxf(v[4 * j + 2] + bias[c0 + 4 * j + 2], 0.f); x.w = fmaxf(v[4 * j + 3] + bias[c0 + 4 * j + 3], 0.f);
                        tf32_split(x.x, h.x, lo4.x); tf32_split(x.y, h.y, lo4.y); tf32_split(x.z, h.z, lo4.z); tf32_split(x.w, h.w, lo4.w);
                        const uint32_t off = umma_off(row, c0 + 4 * j, sbon);
                        *reinterpret_cast<float4 *>(Ahi + off) = h;
                        if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + row, c0 + 4 * j, sbon)) = lo4;
                        else *reinterpret_cast<float4 *>(Alo + off) = lo4;
                        if (mine) *reinterpret_cast<float4 *>(act_row + c0 + 4 * j) = x;      // kept for dW
                        mk |= ((x.x > 0.f ? 1u : 0u) | (x.y > 0.f ? 2u : 0u) | (x.z > 0.f ? 4u : 0u) | (x.w > 0.f ? 8u : 0u)) << (4 * j);
                    }
                    if (!mine) mk = 0u;
                    if (l == 0) hm1 = mk; else if (l == 1) hm2 = mk; else if (l == 2) hm3 = mk; else hm4 = mk;
                }
            } else {
                // head: Q(s, .), loss, dLoss/dHead -> next A operand (K = 32) and the dz scratch
                const uint32_t sbon = umma_sbo(T.N_pad);
                if (half == 0 && live) {
                    float q[32];
                    if (stack) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] = vpre[j];
                        stack_add_lo(q, s_lo, row, 0);
                    } else tmem_ld32_sum(taddr, second, q);
                    TR_TRACE(21);
                    const int nA = tc.n_actions;
#pragma unroll
                    for (int j = 0; j < 32; ++j) q[j] += bias[j];
                    if (DUELING) {
                        float s = 0.f, V = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) { if (j < nA) s += q[j]; if (j == nA) V = q[j]; }
                        const float mean = s / (float)nA;
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] = V + q[j] - mean;
                    }
                    const int act = s_act[row];
                    float qa = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (j == act) qa = q[j];
                    TR_TRACE(22);
                    float gq = 0.f, lterm = 0.f;
                    if (mine) {
                        const float diff = qa - s_y[row];
                        const float wb = a.src.is_w ? a.src.is_w[gb] : 1.f;
                        if (a.src.abs_err) a.src.abs_err[gb] = fabsf(diff);
                        if (a.loss_kind == 0) {                       // MSELoss (BaseTrainer.py:40)
                            lterm = wb * (diff * diff);
                            gq = (2.f * diff * wb) * a.inv_global_b;
                        } else {                                      // SmoothL1Loss(beta = 1)
                            const float ad = fabsf(diff);
                            lterm = wb * (ad < 1.f ? 0.5f * (diff * diff) : ad - 0.5f);
                            gq = (fminf(fmaxf(diff, -1.f), 1.f) * wb) * a.inv_global_b;
                        }
                    }
                    // the warp's 32 loss terms: butterfly sum, ONE shared-memory add per warp (an atomicAdd per lane on the same
                    // word is a 32-deep compare-and-swap chain: 6 k cycles in the round-2 stage trace) -- and a fixed order
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) lterm += __shfl_xor_sync(0xffffffffu, lterm, off);
                    TR_TRACE(23);
                    if (lane == 0) atomicAdd(&s_loss, lterm);
                    TR_TRACE(24);
                    float g[32];
                    const float inv = 1.f / (float)nA;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float gj;
                        if (DUELING) gj = (j < nA) ? gq * ((j == act ? 1.f : 0.f) - inv) : (j == nA ? gq : 0.f);
                        else gj = (j == act) ? gq : 0.f;
                        g[j] = gj;
                    }
                    TR_TRACE(25);
                    float *dz_row = a.dz_buf + (size_t)gb * tc.dz_stride + T.dz_off;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 x = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]), h, lo4;
                        tf32_split(x.x, h.x, lo4.x); tf32_split(x.y, h.y, lo4.y); tf32_split(x.z, h.z, lo4.z); tf32_split(x.w, h.w, lo4.w);
                        const uint32_t off = umma_off(row, 4 * j, sbon);
                        *reinterpret_cast<float4 *>(Ahi + off) = h;
                        if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + row, 4 * j, sbon)) = lo4;
                        else *reinterpret_cast<float4 *>(Alo + off) = lo4;
                        if (mine) *reinterpret_cast<float4 *>(dz_row + 4 * j) = x;
                    }
                    TR_TRACE(26);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            TR_TRACE(11 + l);
        }

        if (fused && tc.train_img_bytes > tc.img_bytes) mbar_wait(&wbar3, 0);                     // transposed blocks (requested at kernel start)
        // ---------------- dX chain: dZ_{l-1} = (dZ_l * W_l) .* (H_l > 0), l = nl-1 .. 1
        for (int l = nl - 1; l >= 1; --l) {
            const TcLayer T = tc.L[l];
            const uint32_t sbo = umma_sbo(T.N_pad);             // reduction runs over this layer's outputs
            const uint32_t dcol = (uint32_t)((l + 1) & 1) * (uint32_t)tc.dstride;     // the head used (nl-1)&1: alternate from there
            const uint32_t second = tc.concat ? (uint32_t)T.K_pad : 0u;
            if (tid == 0) {
                if (stack) issue_3xtf32_stacked(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(W + T.t_hi_off), sbo), kTcTile, T.K_pad, T.N_pad / 8);
                else issue_3xtf32(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(Alo), sbo),
                                  umma_desc(smem_u32(W + T.t_hi_off), sbo), umma_desc(smem_u32(W + T.t_lo_off), sbo), kTcTile, T.K_pad,
                                  T.N_pad / 8, tc.concat != 0);
                umma_commit(&mbar);
            }
            // ReLU'(H_l): the sign mask this thread kept in the forward epilogue (K_pad <= 64: one 32-column chunk per thread)
            const uint32_t hmask = (l == 1) ? hm1 : (l == 2) ? hm2 : (l == 3) ? hm3 : hm4;
            mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + dcol;
            const uint32_t sbon = umma_sbo(T.K_pad);
            float *dz_row = a.dz_buf + (size_t)gb * tc.dz_stride + tc.L[l - 1].dz_off;
            float vpre[32];                                      // this warp's accumulator chunk, loaded while the lo warps park theirs
            if (stack) {                                           // lo*hi block (rows [R, 2R)) -> scratch
                if (quad * 32 >= R && quad * 32 < 2 * R)
                    for (int c0 = half * 32; c0 < T.K_pad && c0 < 64; c0 += 64) stack_park_lo(taddr, c0, s_lo, row - R);
                if (live && half * 32 < T.K_pad) tmem_ld32_sum(taddr + (uint32_t)(half * 32), second, vpre);
                __syncthreads();
            }
            {
                const int c0 = half * 32;
                if (live && c0 < T.K_pad) {
                    float v[32];
                    if (stack) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = vpre[j];
                        stack_add_lo(v, s_lo, row, c0);
                    } else tmem_ld32_sum(taddr + (uint32_t)c0, second, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t m4 = hmask >> (4 * j);
                        float4 x, h, lo4;
                        x.x = (m4 & 1u) ? v[4 * j + 0] : 0.f; x.y = (m4 & 2u) ? v[4 * j + 1] : 0.f;
                        x.z = (m4 & 4u) ? v[4 * j + 2] : 0.f; x.w = (m4 & 8u) ? v[4 * j + 3] : 0.f;
                        if (mine) *reinterpret_cast<float4 *>(dz_row + c0 + 4 * j) = x;
                        if (l > 1) {
                            tf32_split(x.x, h.x, lo4.x); tf32_split(x.y, h.y, lo4.y); tf32_split(x.z, h.z, lo4.z); tf32_split(x.w, h.w, lo4.w);
                            const uint32_t off = umma_off(row, c0 + 4 * j, sbon);
                            *reinterpret_cast<float4 *>(Ahi + off) = h;
                            if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + row, c0 + 4 * j, sbon)) = lo4;
                            else *reinterpret_cast<float4 *>(Alo + off) = lo4;
                        }
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            TR_TRACE(16 + l);
        }
    }
    TR_TRACE(20);
    if (tid == 0) a.loss_partials[blockIdx.x] = s_loss;
    tc_fence_before();
    __syncthreads();
    if (warp == kCtl / 32) { tc_fence_after(); tmem_dealloc(tmem_base_s, (uint32_t)tc.tmem_cols); }
}

// ------------------------------------------------------------------ split-K weight gradients
constexpr uint32_t kDwBlk = (kDwChunk / 4) * kMnAtom;      // bytes of one 32-feature block: 128 samples x 128 B = 16 KB (the LBO)
constexpr int kDwABlocks = 4;                               // A operand: M = 128 = 4 feature blocks (input features + the ones column)

// rows [128 samples][width floats] in global memory -> hi/lo MN-major operand blocks.  8 consecutive lanes copy one 128-byte
// row segment (coalesced), each as ONE 16-byte store into the swizzled position: no transposition, no bank conflicts.
// LOG2F4 = log2(float4 per row) (5 for the 128-wide A operand, 3 / 4 for a 32- / 64-wide B operand); U float4 per thread,
// ALL loaded before the first is converted (the loads are the latency that matters: one round trip per operand).
// ones_col >= 0: that feature column is set to 1 for valid samples (bias gradient).
template <int LOG2F4, int U>
__device__ __forceinline__ void dw_load_rows(const float *const *rows, int width, float4 (&v)[U])
{
    // nothing but loads here: a register that a load is still going to write must not be touched again before the data
    // is consumed, or the warp stalls on that load before issuing the next one (the ones column is applied at store time)
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = threadIdx.x + u * kTcThreads;
        const int b = i >> LOG2F4, jc = i & ((1 << LOG2F4) - 1);
        const float *r = (b < kDwChunk) ? rows[b] : nullptr;
        v[u] = (r && 4 * jc < width) ? __ldg(reinterpret_cast<const float4 *>(r) + jc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// ones_col >= 0: that feature column (a multiple of 4: K_real % 4 == 0, tc_train_init) becomes 1 for valid samples
template <int LOG2F4, int U>
__device__ __forceinline__ void dw_store_rows(const float4 (&v)[U], const float *const *rows, int ones_col, unsigned char *hi, unsigned char *lo)
{
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = threadIdx.x + u * kTcThreads;
        const int b = i >> LOG2F4, jc = i & ((1 << LOG2F4) - 1);
        if (b >= kDwChunk) continue;
        float4 x = v[u];
        if (ones_col >= 0 && (ones_col >> 2) == jc && rows[b]) x.x = 1.f;
        float4 h, l;
        tf32_split(x.x, h.x, l.x); tf32_split(x.y, h.y, l.y); tf32_split(x.z, h.z, l.z); tf32_split(x.w, h.w, l.w);
        const uint32_t off = umma_mn_off(4 * jc, b, kDwBlk);
        *reinterpret_cast<float4 *>(hi + off) = h;
        *reinterpret_cast<float4 *>(lo + off) = l;
    }
}

template <bool FUSE_ADAM>
__global__ void __launch_bounds__(kTcThreads, 1) tc_dw_kernel(TcNet tc, TcDwArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int l = blockIdx.x % tc.n_layers, slice = blockIdx.x / tc.n_layers;
    const int chunk = slice;                                  // partial index
    const TcLayer T = tc.L[l];
    const int rowsA = T.K_real + 1;                           // input features + the all-ones column (bias gradient)
    const int nbB = T.N_pad / 32;                             // 32-wide output blocks of dZ (N_pad is 32 or 64)
    // A = [act ; 1] as [sample][feature], 4 feature blocks (M = 128; columns past rowsA are zero), hi then lo;
    // B = dZ as [sample][out], hi blocks then lo blocks (adjacent: the concatenated 3xTF32 product reads them as N = 2 N_pad)
    unsigned char *Ahi = smem, *Alo = Ahi + kDwABlocks * kDwBlk, *Bhi = Alo + kDwABlocks * kDwBlk, *Blo = Bhi + nbB * kDwBlk;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ const float *rows[2][kDwChunk];               // double buffered: chunk c + 1 is resolved while chunk c is loaded
    __shared__ const float *drows[2][kDwChunk];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
    DW_TRACE(0);
    // the last warp allocates TMEM and initialises the barrier while warps 0-3 resolve the first chunk's rows (tc_forward.cu)
    constexpr int kCtl = kTcThreads - 32;
    if (tid == kCtl) { mbar_init(&mbar, 1); fence_barrier_init(); }
    if (warp == kCtl / 32) { __syncwarp(); tmem_alloc(&tmem_base_s, (uint32_t)tc.dstride); }
    uint32_t pkey[4];
    Philox::gen(a.src.key, a.src.epoch, 0x5A17ull, pkey);
    auto resolve_chunk = [&](int c, int buf) {                // row pointers of chunk c (128 samples)
        if (tid < kDwChunk) {
            const int b = c * kDwChunk + tid;
            const float *p = nullptr, *dzp = nullptr;
            if (c < a.n_chunks && b < a.B) {
                p = (l == 0) ? resolve_transition(a.src, b, tc.in_dim, pkey).s : a.act_buf + (size_t)b * tc.act_stride + T.act_off;
                dzp = a.dz_buf + (size_t)b * tc.dz_stride + T.dz_off;
            }
            rows[buf][tid] = p; drows[buf][tid] = dzp;
        }
    };
    resolve_chunk(slice, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // PDL: hidden activations and dZ come from the training chain (the predecessor); the layer-0 CTAs' A operand is
    // built from replay rows (written >= 2 kernels back) and is gathered before the wait
    if (l != 0) { pdl_wait(); pdl_trigger(); }
    DW_TRACE(1);
    // A: 128 samples x 128 columns (features, the ones column, zero padding) = 16 float4 per thread; B: 128 x N_pad.
    // Persistent over this slice's chunks: the loads of chunk c + 1 are in flight while the tensor core works on chunk c,
    // whose products accumulate in the same TMEM columns -> one partial per slice however large the batch.
    float4 va[16], vb[8];
    auto load_chunk = [&](int buf) {
        dw_load_rows<5, 16>(rows[buf], T.K_real, va);
        if (T.N_pad == 32) {
#pragma unroll
            for (int u = 4; u < 8; ++u) vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            dw_load_rows<3, 4>(drows[buf], 32, reinterpret_cast<float4 (&)[4]>(vb));
        } else {
            dw_load_rows<4, 8>(drows[buf], 64, vb);         // N_pad = 64: two 32-column blocks, kDwBlk apart
        }
    };
    if (l == 0) {                                              // replay rows first, dZ after the predecessor has finished
        dw_load_rows<5, 16>(rows[0], T.K_real, va);
        pdl_wait(); pdl_trigger();
        if (T.N_pad == 32) {
#pragma unroll
            for (int u = 4; u < 8; ++u) vb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            dw_load_rows<3, 4>(drows[0], 32, reinterpret_cast<float4 (&)[4]>(vb));
        } else {
            dw_load_rows<4, 8>(drows[0], 64, vb);
        }
    } else {
        load_chunk(0);
    }
    DW_TRACE(2);
    uint32_t mphase = 0;
    int it = 0;
    for (int c = slice; c < a.n_chunks; c += a.n_slices, ++it) {
        if (it > 0) { mbar_wait(&mbar, mphase); mphase ^= 1; tc_fence_after(); }     // the MMAs of the previous chunk have read SMEM
        dw_store_rows<5, 16>(va, rows[it & 1], T.K_real, Ahi, Alo);
        if (T.N_pad == 32) dw_store_rows<3, 4>(reinterpret_cast<float4 (&)[4]>(vb), drows[it & 1], -1, Bhi, Blo);
        else dw_store_rows<4, 8>(vb, drows[it & 1], -1, Bhi, Blo);
        const int cn = c + a.n_slices;
        resolve_chunk(cn, (it + 1) & 1);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        DW_TRACE(4);
        if (tid == 0) {
            issue_3xtf32_mn(tmem, umma_desc_mn(smem_u32(Ahi), kDwBlk), umma_desc_mn(smem_u32(Alo), kDwBlk), umma_desc_mn(smem_u32(Bhi), kDwBlk),
                            umma_desc_mn(smem_u32(Blo), kDwBlk), kTcTile, T.N_pad, kDwChunk / 8, tc.concat != 0, it > 0 ? 1u : 0u);
            umma_commit(&mbar);
        }
        if (cn < a.n_chunks) load_chunk((it + 1) & 1);         // next chunk's rows -> registers while the MMAs run
    }
    DW_TRACE(5);
    mbar_wait(&mbar, mphase);
    tc_fence_after();
    DW_TRACE(6);
    // epilogue: accumulator row f = input feature (or the ones column), column o = output unit.  Lanes hold consecutive f:
    // every store instruction writes 32 consecutive floats of one weight row.
    // Partial slice `chunk`: plain floats for the separate optimiser kernel, or 8-byte words {epoch : value} for the fused tail.
    // Lanes hold consecutive f, so every store instruction writes 32 consecutive elements of one weight row; the 32 columns of a
    // thread walk the rows with a pointer increment and a predicate each (the address arithmetic used to dominate this epilogue).
    float *part = a.partials + (size_t)chunk * a.P;
    unsigned long long *part64 = FUSE_ADAM ? a.part64 + (size_t)chunk * a.P : nullptr;
    const unsigned long long tag = (unsigned long long)a.ll_epoch << 32;
    const int f = quad * 32 + lane;
    for (int c0 = half * 32; c0 < T.N_pad; c0 += 64) {
        if (quad * 32 >= rowsA) break;
        float v[32];
        tmem_ld32_sum(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, tc.concat ? (uint32_t)T.N_pad : 0u, v);
        const int n_main = min(max(T.out_main - c0, 0), 32), n_all = min(max(T.N_real - c0, 0), 32);   // columns [0, n_main): main block, [n_main, n_all): value head
        if (f > T.K_real) continue;
        const bool brow = (f == T.K_real);                       // the ones column: bias gradients (stride 1), else weight row f (stride K_real)
        const int stride = brow ? 1 : T.K_real;
        const int i_main = brow ? T.b_off + c0 : T.w_off + f + c0 * T.K_real;
        const int i_val = brow ? T.b2_off + (c0 - T.out_main) : (T.w2_off >= 0 ? T.w2_off : 0) + f + (c0 - T.out_main) * T.K_real;
        if (FUSE_ADAM) {
            unsigned long long *p = part64 + i_main;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < n_main) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(tag | __float_as_uint(v[j])) : "memory");
                p += stride;
            }
            if (n_all > n_main) {
                unsigned long long *q = part64 + i_val;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= n_main && j < n_all) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(q), "l"(tag | __float_as_uint(v[j])) : "memory");
                    q += stride;
                }
            }
        } else {
            float *p = part + i_main;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < n_main) *p = v[j];
                p += stride;
            }
            if (n_all > n_main) {
                float *q = part + i_val;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= n_main && j < n_all) *q = v[j];
                    q += stride;
                }
            }
        }
    }
    DW_TRACE(7);
    tc_fence_before();
    __syncthreads();
    if (warp == kCtl / 32) { tc_fence_after(); tmem_dealloc(tmem, (uint32_t)tc.dstride); }
    DW_TRACE(8);
    if (!FUSE_ADAM) return;
    // ---- fused optimiser tail: this CTA reduces its slice of the parameter vector over all partials in reduce_adam_kernel's
    // order and applies Adam.  Every word it needs is polled until it carries this launch's epoch (all CTAs are resident: grid
    // <= SMs, one CTA per SM, and each writes its partial before it polls) -- nothing to fence, no barrier to wait at.
    const int per = (a.P + (int)gridDim.x - 1) / (int)gridDim.x;
    const int i_end = min(a.P, ((int)blockIdx.x + 1) * per);
    for (int i = (int)blockIdx.x * per + tid; i < i_end; i += kTcThreads) {
        const unsigned long long *pp = a.part64 + i;
        const uint32_t ep = a.ll_epoch;
        auto poll = [ep](const unsigned long long *q) {
            unsigned long long w;
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(q) : "memory");
            return w;
        };
        float g4[4];
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
            float acc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = 0.f;
            int c = cg;
            for (; c + 28 < a.adam.nparts; c += 32) {
                unsigned long long w[8];
                bool ok;
                do {
                    ok = true;
#pragma unroll
                    for (int u = 0; u < 8; ++u) w[u] = poll(pp + (size_t)(c + 4 * u) * a.P);
#pragma unroll
                    for (int u = 0; u < 8; ++u) ok = ok && (uint32_t)(w[u] >> 32) == ep;
                } while (!ok);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc[u] += __uint_as_float((uint32_t)w[u]);
            }
            for (; c < a.adam.nparts; c += 4) {
                unsigned long long w;
                do { w = poll(pp + (size_t)c * a.P); } while ((uint32_t)(w >> 32) != ep);
                acc[0] += __uint_as_float((uint32_t)w);
            }
            g4[cg] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
        }
        const float g = (g4[0] + g4[1]) + (g4[2] + g4[3]);
        a.ptrs.grad[i] = g;
        adam_update_one(a.adam, a.ptrs, i, g);
    }
    if (blockIdx.x == 0 && warp == 7 && a.ptrs.loss_out) {          // loss = sum of the training kernel's per-CTA partials / B
        float sl = 0.f;
        for (int c = lane; c < a.adam.n_loss_parts; c += 32) sl += __ldcg(a.ptrs.loss_partials + c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sl += __shfl_xor_sync(0xffffffffu, sl, off);
        if (lane == 0) *a.ptrs.loss_out = sl * a.adam.inv_b;
    }
}

typedef void (*TrainKernel)(TcNet, TcTrainArgs);
template <bool S, int N>
static TrainKernel pick_train_d(bool dueling) { return dueling ? tc_train_kernel<S, N, true> : tc_train_kernel<S, N, false>; }
template <bool S>
static TrainKernel pick_train_n(int npre, bool dueling) { return npre == 0 ? pick_train_d<S, 0>(dueling) : npre == 1 ? pick_train_d<S, 1>(dueling) : pick_train_d<S, 2>(dueling); }
static TrainKernel pick_train_kernel(bool stack, int npre, bool dueling) { return stack ? pick_train_n<true>(npre, dueling) : pick_train_n<false>(npre, dueling); }

static size_t train_smem_bytes(const TcNet &tc, int R)
{
    return (size_t)2 * (R / 8) * umma_sbo(tc.max_k) + (size_t)tc.train_img_bytes + (size_t)R * kLoLd * 4;   // + the stacked-3xTF32 scratch
}
static size_t dw_smem_bytes(const TcNet &tc)
{
    int maxN = 32;
    for (int l = 0; l < tc.n_layers; ++l) if (tc.L[l].N_pad > maxN) maxN = tc.L[l].N_pad;
    return (size_t)2 * (kDwABlocks + maxN / 32) * kDwBlk;       // A hi/lo (4 blocks each) + B hi/lo
}

int tc_train_init(uavrl_learner *l)
{
    l->tc_train_ok = false;
    const TcNet &tc = l->tc;
    for (int i = 0; i < tc.n_layers; ++i)
        if (tc.L[i].K_real + 1 > 128 || tc.L[i].K_real % 4 != 0 || (tc.L[i].N_pad != 32 && tc.L[i].N_pad != 64)) return 0;   // ones column / float4 chunks / 1-2 blocks
    // the M=128 MMA reads 16 row groups from each A buffer: with fewer real rows it runs into the next buffers,
    // which must still be inside the CTA's allocation
    if (train_smem_bytes(tc, 32) > 227 * 1024 || dw_smem_bytes(tc) > 227 * 1024) return 0;
    if (train_smem_bytes(tc, 32) < (size_t)(32 / 8) * umma_sbo(tc.max_k) + (size_t)16 * umma_sbo(tc.max_k)) return 0;
    for (int st = 0; st < 2; ++st)
        for (int np = 0; np < 3; ++np)
            for (int du = 0; du < 2; ++du)
                UAVRL_CUDA(cudaFuncSetAttribute(pick_train_kernel(st != 0, np, du != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(train_smem_bytes(tc, 64) <= 227 * 1024 ? train_smem_bytes(tc, 64) : train_smem_bytes(tc, 32))));
    UAVRL_CUDA(cudaFuncSetAttribute(tc_dw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem_bytes(tc)));
    UAVRL_CUDA(cudaFuncSetAttribute(tc_dw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dw_smem_bytes(tc)));
    const size_t cap = (size_t)l->cfg.batch_size;
    UAVRL_CUDA(cudaMalloc((void **)&l->act_buf, cap * (size_t)(tc.act_stride > 0 ? tc.act_stride : 4) * 4));
    UAVRL_CUDA(cudaMalloc((void **)&l->dz_buf, cap * (size_t)tc.dz_stride * 4));
    l->train_cap = (int32_t)cap;
    l->tc_train_ok = true;
    return 0;
}

std::atomic<int> g_fuse_td{1};               // uavrl_set_fuse_td(); default on
// rows per tile of the training kernel: 32 while the batch fits one wave of 32-row tiles, else 64 (when the 64-row operands fit)
static int train_rows_per_tile(const TcNet &tc, int B)
{
    return (B > 32 * 148 && train_smem_bytes(tc, 64) <= 227 * 1024) ? 64 : 32;
}
bool tc_train_can_fuse_td(const uavrl_learner *l, int B)
{
    // one tile per CTA (32-row tiles up to 4 736 samples, 64-row tiles up to 9 472): the weight images are restaged inside the
    // kernel, which only pays when a CTA does it once; larger batches keep the separate TD-target kernel(s) whose CTAs reuse
    // one image over several tiles
    if (!g_fuse_td.load() || !l->tc_train_ok) return false;
    const int R = train_rows_per_tile(l->tc, B);
    return (B + R - 1) / R <= 148;
}
std::atomic<int> g_fuse_dw_adam{0};          // uavrl_set_fuse_dw_adam(); default off: measured no faster than the PDL-chained pair

int launch_tc_train(uavrl_learner *l, const BatchSrc &src, int B, int global_batch, const float *y, int *n_grad_parts,
                    int *n_loss_parts, cudaStream_t st, cudaEvent_t after_chain, const AdamArgs *adam, float *loss_out, bool *adam_done,
                    bool fused_td)
{
    const TcNet &tc = l->tc;
    TcTrainArgs a;
    memset(&a, 0, sizeof(a));
    a.img = l->tc_img_local; a.src = src; a.B = B; a.y = y; a.inv_global_b = 1.0f / (float)global_batch;
    a.act_buf = l->act_buf; a.dz_buf = l->dz_buf; a.loss_partials = l->loss_partials;
    a.loss_kind = l->cfg.loss_kind;
    a.fused_td = fused_td ? 1 : 0; a.algo = l->cfg.algo; a.gamma = l->cfg.gamma; a.img_target = l->tc_img_target;
    a.R = train_rows_per_tile(tc, B);
    a.n_tiles = (B + a.R - 1) / a.R;
    const int grid = a.n_tiles < 148 ? a.n_tiles : 148;
    const bool chain = l->pdl_chain && g_pdl.load();
    if (fused_td && a.n_tiles > grid) return fail(UAVRL_ERR_INVALID, "fused TD needs one tile per CTA");
    static const bool trace_on = getenv("UAVRL_TC_TRACE") != nullptr;
    long long *tr = nullptr;
    if (trace_on) { UAVRL_CUDA(cudaMalloc((void **)&tr, 48 * sizeof(long long))); UAVRL_CUDA(cudaMemset(tr, 0, 48 * sizeof(long long))); a.trace = tr + 16; }
    const bool use_pdl = chain && (fused_td ? (l->pdl_prev == kPdlEnv) : (l->pdl_prev == kPdlTd));
    const bool stack = tc.concat != 0 && tc.dstride <= 128;
    const int npre = fused_td ? (l->cfg.algo != UAVRL_ALGO_DQN ? 2 : 1) : 0;
    auto train_fn = pick_train_kernel(stack, npre, tc.dueling != 0);
    UAVRL_CUDA(launch_kernel(train_fn, dim3(grid), dim3(kTcThreads), train_smem_bytes(tc, a.R), st, use_pdl, tc, a));
    // experiment (with UAVRL_TC_TRACE): the same launch again, back to back -- the kernel is idempotent, the second run finds
    // its code in the instruction caches, and the stage trace printed below is the second run's
    static const bool twice = trace_on && getenv("UAVRL_TRAIN_TWICE") != nullptr;
    if (twice) UAVRL_CUDA(launch_kernel(train_fn, dim3(grid), dim3(kTcThreads), train_smem_bytes(tc, a.R), st, false, tc, a));
    l->pdl_prev = chain ? kPdlTrain : kPdlNone;
    UAVRL_LAUNCHED();
    if (after_chain) UAVRL_CUDA(cudaEventRecord(after_chain, st));
    TcDwArgs d;
    memset(&d, 0, sizeof(d));
    d.src = src; d.B = B; d.n_chunks = (B + kDwChunk - 1) / kDwChunk; d.P = l->net.P;
    d.act_buf = l->act_buf; d.dz_buf = l->dz_buf; d.partials = l->partials;
    static int n_sm = 0;
    if (n_sm == 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    // one CTA per SM (192 KB of shared memory): at most n_sm / n_layers slices per layer, each looping over its chunks
    const int max_slices = n_sm / tc.n_layers > 0 ? n_sm / tc.n_layers : 1;
    d.n_slices = d.n_chunks < max_slices ? d.n_chunks : max_slices;
    const int dw_grid = d.n_slices * tc.n_layers;
    // Fused optimiser tail: the kernel's grid barrier needs every CTA resident at once -- one CTA per SM (192 KB of shared
    // memory each), so only when the grid fits the SMs; larger batches keep the separate reduce_adam_kernel.
    const bool fuse = adam != nullptr && g_fuse_dw_adam.load() && dw_grid <= n_sm;
    if (adam_done) *adam_done = fuse;
    if (fuse) {
        if (!l->dw_bar) {                                        // the {epoch : value} partial buffer [max_slices][P], zeroed once
            const size_t n = (size_t)max_slices * (size_t)l->net.P * sizeof(unsigned long long);
            UAVRL_CUDA(cudaMalloc((void **)&l->dw_bar, n));
            UAVRL_CUDA(cudaMemsetAsync(l->dw_bar, 0, n, st));
            l->dw_bar_total = 0;
        }
        l->dw_bar_total += 1;
        if ((uint32_t)l->dw_bar_total == 0u) l->dw_bar_total += 1;     // tag 0 = "never written"
        d.fuse_adam = 1; d.part64 = l->dw_bar; d.ll_epoch = (uint32_t)l->dw_bar_total;
        d.adam = *adam; d.adam.nparts = d.n_slices; d.adam.n_loss_parts = grid;
        AdamPtrs &q = d.ptrs;
        q.partials = l->partials; q.loss_partials = l->loss_partials; q.grad = l->grad; q.local = l->local; q.m = l->m; q.v = l->v;
        q.target = l->target; q.img_local = l->img_local; q.img_target = l->img_target; q.img_map = l->img_map;
        q.tc_local = (float *)l->tc_img_local; q.tc_target = (float *)l->tc_img_target; q.tc_hi = l->tc_hi_map; q.tc_lo = l->tc_lo_map;
        q.tc_hi2 = l->tc_hi2_map; q.tc_lo2 = l->tc_lo2_map; q.loss_out = loss_out;
    }
    if (trace_on) d.trace = tr;
    UAVRL_CUDA(launch_kernel(fuse ? tc_dw_kernel<true> : tc_dw_kernel<false>, dim3(dw_grid), dim3(kTcThreads), dw_smem_bytes(tc), st,
                             chain && !after_chain, tc, d));
    l->pdl_prev = chain ? (fuse ? kPdlAdam : kPdlDw) : kPdlNone;
    UAVRL_LAUNCHED();
    if (trace_on) {
        long long h[48];
        UAVRL_CUDA(cudaStreamSynchronize(st));
        UAVRL_CUDA(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(tr);
        fprintf(stderr, "[train_trace] B=%d R=%d fused_td=%d cycles since start:", B, a.R, a.fused_td);
        for (int i = 1; i < 32; ++i) if (h[16 + i]) fprintf(stderr, " [%d]=%lld", i, h[16 + i] - h[16]);
        fprintf(stderr, "\n");
        fprintf(stderr, "[dw_trace] B=%d chunks=%d (CTA 0 = layer 0) cycles since start:", B, d.n_chunks);
        for (int i = 1; i < 9; ++i) fprintf(stderr, " [%d]=%lld", i, h[i] - h[0]);
        fprintf(stderr, "\n");
    }
    *n_grad_parts = d.n_slices;
    *n_loss_parts = grid;
    return 0;
}

}  // namespace uavrl

extern "C" int uavrl_set_fuse_dw_adam(int32_t on) { uavrl::g_fuse_dw_adam.store(on ? 1 : 0); return 0; }
extern "C" int uavrl_set_fuse_td(int32_t on) { uavrl::g_fuse_td.store(on ? 1 : 0); return 0; }
extern "C" int uavrl_learner_td_fused(const uavrl_learner *l, int32_t batch) { return (l && l->tc_ok && l->use_tc && uavrl::tc_train_can_fuse_td(l, batch)) ? 1 : 0; }
