// tc_forward.cu -- the Q-network forward chain on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Used for Trainer.get_action (Trainer/DuelingDQN_Trainer.py:86-97) and for the TD-target part of
// Trainer.update (target-network forward, :164-171; DQN_Trainer.py:109; DDQN_Trainer.py:94-95).
//
// One CTA = one tile of 128 samples (UMMA M = 128).  The whole network lives in SMEM as B operands in the
// canonical K-major layout, pre-split into TF32 hi/lo images that the optimiser kernel keeps current
// (one TMA bulk copy stages it).  Per layer:   D[128 x N] (TMEM, fp32) = A[128 x K] * W[N x K]^T
// issued by ONE thread as 3 x K/8 `tcgen05.mma.kind::tf32` (hi*hi + hi*lo + lo*hi: the 3xTF32 split,
// measured 1.4e-6 relative -- fp32 grade, so results stay inside the parity tolerance of the fp32 path).
// The epilogue reads the accumulator with `tcgen05.ld`, adds the bias, applies ReLU, re-splits and writes
// the next layer's A operand straight back into SMEM -- activations never touch HBM.  The head's epilogue
// forms Q (dueling combine), then eps-greedy / argmax / max / gather depending on the mode.
#include "tc_forward.cuh"
#include "env_block.cuh"

#include <stdlib.h>
#include <string.h>

#include "tma.cuh"
#include "umma.cuh"

namespace uavrl {

static inline int rup(int x, int m) { return (x + m - 1) / m * m; }

int tc_build(const uavrl_learner_config &c, const NetDev &net, TcNet &tc, std::vector<int32_t> &hi_map, std::vector<int32_t> &lo_map,
             std::vector<int32_t> &hi2_map, std::vector<int32_t> &lo2_map)
{
    memset(&tc, 0, sizeof(tc));
    tc.n_layers = net.n_layers; tc.in_dim = net.in_dim; tc.n_actions = net.n_actions; tc.dueling = net.dueling;
    if (net.in_dim % 4 != 0) return -1;
    int off = 0, maxK = 0, k_pad = rup(net.in_dim, 8);
    for (int l = 0; l < net.n_layers; ++l) {
        const LayerDev &L = net.L[l];
        TcLayer &T = tc.L[l];
        const bool head = (l == net.n_layers - 1);
        T.K_real = L.in; T.K_pad = k_pad; T.N_real = L.out;
        T.N_pad = head ? 32 : rup(L.out, 16);
        if (T.N_pad > 128 || T.K_pad > 128 || (head && L.out > 32)) return -1;
        const int bytes = T.N_pad * T.K_pad * 4;
        T.hi_off = off; off += bytes;
        T.lo_off = off; off += bytes;
        if (T.K_pad > maxK) maxK = T.K_pad;
        T.w_off = L.w_off; T.b_off = L.b_off; T.w2_off = L.w2_off; T.b2_off = L.b2_off;
        T.out_main = L.out_main;
        k_pad = T.N_pad;
    }
    tc.bias_base = off;
    int boff = 0;
    for (int l = 0; l < net.n_layers; ++l) { tc.L[l].bias_off = boff; boff += tc.L[l].N_pad; }
    tc.img_bytes = rup(off + boff * 4, 16);
    // transposed blocks for the dX chain (layers >= 1): B' = W^T, rows = input units (K_pad), reduction = outputs (N_pad)
    int toff = tc.img_bytes;
    for (int l = 0; l < net.n_layers; ++l) {
        TcLayer &T = tc.L[l];
        T.t_hi_off = T.t_lo_off = -1;
        if (l == 0) continue;
        const int bytes = T.K_pad * T.N_pad * 4;
        T.t_hi_off = toff; toff += bytes;
        T.t_lo_off = toff; toff += bytes;
    }
    tc.train_img_bytes = rup(toff, 16);
    tc.a_bytes = (int)umma_tile_bytes(kTcTile, maxK);
    tc.max_k = maxK;
    {   // accumulator columns: N_pad per product block; the concatenated scheme keeps hi*hi+lo*hi and hi*lo side by side
        static const bool three_pass = getenv("UAVRL_TC_3PASS") != nullptr;
        int maxN = 32;
        for (int l = 0; l < net.n_layers; ++l) if (tc.L[l].N_pad > maxN) maxN = tc.L[l].N_pad;
        tc.concat = three_pass ? 0 : 1;
        const int need = (tc.concat ? 2 : 1) * maxN;
        tc.dstride = need <= 64 ? 64 : need <= 128 ? 128 : 256;
        tc.tmem_cols = 2 * tc.dstride;
    }
    // per-sample scratch: act = inputs of layers 1.. (K_pad each), dz = output derivatives of every layer (N_pad each)
    int ao = 0, dzo = 0;
    for (int l = 0; l < net.n_layers; ++l) {
        tc.L[l].act_off = (l == 0) ? -1 : ao;
        if (l > 0) ao += tc.L[l].K_pad;
        tc.L[l].dz_off = dzo; dzo += tc.L[l].N_pad;
    }
    tc.act_stride = ao; tc.dz_stride = dzo;
    // parameter -> image maps (float indices)
    hi_map.assign((size_t)net.P, -1); lo_map.assign((size_t)net.P, -1);
    hi2_map.assign((size_t)net.P, -1); lo2_map.assign((size_t)net.P, -1);
    auto elem = [](int k_pad_cols, int base, int n, int k) {     // float index of element (row n, col k), K-major canonical
        const int sbo = (k_pad_cols / 4) * 128;
        return (base + (n >> 3) * sbo + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4) / 4;
    };
    for (int l = 0; l < net.n_layers; ++l) {
        const LayerDev &L = net.L[l];
        const TcLayer &T = tc.L[l];
        const int out_main = T.out_main;
        for (int o = 0; o < L.out; ++o) {
            const bool vrow = (o >= out_main);                   // dueling value row
            for (int k = 0; k < L.in; ++k) {
                const size_t pi = vrow ? (size_t)L.w2_off + (size_t)(o - out_main) * L.in + k : (size_t)L.w_off + (size_t)o * L.in + k;
                hi_map[pi] = elem(T.K_pad, T.hi_off, o, k);
                lo_map[pi] = elem(T.K_pad, T.lo_off, o, k);
                if (l > 0) {                                      // W^T: row k, column o
                    hi2_map[pi] = elem(T.N_pad, T.t_hi_off, k, o);
                    lo2_map[pi] = elem(T.N_pad, T.t_lo_off, k, o);
                }
            }
            hi_map[vrow ? (size_t)L.b2_off + (o - out_main) : (size_t)L.b_off + o] = tc.bias_base / 4 + T.bias_off + o;
        }
    }
    (void)c;
    return 0;
}

size_t tc_smem_bytes(const TcNet &tc) { return (size_t)2 * tc.a_bytes + (size_t)tc.img_bytes; }

#define TC_TRACE(slot) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[slot] = clock64(); } while (0)

// FUSE_ENV: after a tile's actions are written, the same CTA steps those R envs (env_block.cuh) -- the act -> step
// dependency is per env, so no grid-wide boundary is needed between Trainer.get_action and BaseEnv.Move_Agent.
// ACT (a.mode == kTcAct), STACK (R <= 64 on the concatenated scheme) and DUELING are compile-time: the kernel a pass runs carries
// no code of the other modes (tc_train.cu: instruction fetch was a third of the live warps' stalls in the generic kernels)
template <bool FUSE_ENV, bool ACT, bool STACK, bool DUELING>
__global__ void __launch_bounds__(kTcThreads, 1) tc_forward_kernel_t(TcNet tc, TcArgs a, EnvFuse ef)
{
    TC_TRACE(0);
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *Ahi = smem, *Alo = smem + tc.a_bytes, *W = smem + 2 * tc.a_bytes;
    __shared__ uint64_t wbar, wbar2, mbar;                   // weights of layer 0 | every other layer + the biases | MMA completion
    __shared__ uint32_t tmem_base_s;
    __shared__ const float *rows[kTcTile];
    __shared__ float s_rew[kTcTile], s_done[kTcTile];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3, half = warp >> 2;
    // The last warp is the control warp: it initialises the two mbarriers, issues the weight copies and allocates TMEM while
    // the other warps already gather the first tile -- nobody needs TMEM or the barriers before the first MMA, so the CTA-wide
    // barrier that publishes them is the one in front of the first weight wait (800 cycles of allocation off the chain).
    constexpr int kCtl = kTcThreads - 32;
    const bool early_w = (a.pdl & kPdlEarlyWeights) != 0;
    // The image travels in two pieces: layer 0's hi|lo block first (all the first MMA needs), the other layers and the biases
    // behind it on a second barrier that is first waited for in layer 0's epilogue -- when the image can only be requested
    // after the dependent-launch wait (act after the optimiser step) half of its latency is covered by the first layer.
    const uint32_t w_split = tc.n_layers > 1 ? (uint32_t)tc.L[1].hi_off : (uint32_t)tc.img_bytes;
    auto stage_image = [&]() {
        fence_proxy_async();
        bulk_g2s_chunked(W, a.img, w_split, &wbar);
        if (w_split < (uint32_t)tc.img_bytes) bulk_g2s_chunked(W + w_split, a.img + w_split, (uint32_t)tc.img_bytes - w_split, &wbar2);
    };
    if (tid == kCtl) {
        mbar_init(&wbar, 1); mbar_init(&wbar2, 1); mbar_init(&mbar, 1); fence_barrier_init();
        if (early_w) stage_image();
    }
    if (warp == kCtl / 32) { __syncwarp(); tmem_alloc(&tmem_base_s, (uint32_t)tc.tmem_cols); tc_fence_before(); }
    uint32_t tmem = 0;
    TC_TRACE(1);
    // PDL prologue (common.cuh): the weight image may be fetched before the wait when the predecessor does not write
    // it (TD passes after the env step); the first tile's rows may be gathered before the wait when the predecessor
    // does not write them (act after the optimiser kernel: observations were written two kernels back).
    bool waited = (a.pdl & kPdlEarlyRows) == 0;
    if (waited) {
        pdl_wait();
        pdl_trigger();
        if (tid == kCtl && !early_w) stage_image();
    }
    const float *bias_all = reinterpret_cast<const float *>(W + tc.bias_base);

    // act mode: row b of the tile is simply obs[base + b] -- no replay sampling (Philox), no pointer table, no barrier
    constexpr bool direct = ACT;
    uint32_t pkey[4] = {0u, 0u, 0u, 0u};
    if (!direct) Philox::gen(a.src.key, a.src.epoch, 0x5A17ull, pkey);
    uint32_t mphase = 0;
    bool wready = false, w2ready = false;

    // R = real rows per tile (32 / 64 / 128).  The MMA is always M = 128; accumulator rows >= R hold garbage
    // computed from whatever SMEM follows the R-row operand (still inside this CTA's allocation) and are
    // never read.  Small batches use R = 32 so that 4096 samples spread over 128 CTAs instead of 32.
    const int R = a.rows_per_tile, lgR = 31 - __clz(R);
    // R <= 64: stacked 3xTF32 (umma.cuh) -- A_lo lives in rows [R, 2R) of the Ahi buffer, the Alo buffer is the scratch through
    // which the lo*hi block reaches the epilogue warps
    constexpr bool stack = STACK;                                       // = R <= 64 && tc.concat && tc.dstride <= 128 (launch_tc_forward)
    float *s_lo = reinterpret_cast<float *>(Alo);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const int base = tile * R;
        if (!direct) {
            if (tid < R) {
                const int b = base + tid;
                const float *p = nullptr;
                float r = 0.f, d = 0.f;
                if (b < a.n) {
                    const Transition t = resolve_transition(a.src, b, tc.in_dim, pkey);
                    p = a.use_next ? t.s2 : t.s; r = t.r; d = t.d;
                }
                rows[tid] = p; s_rew[tid] = r; s_done[tid] = d;
            }
            __syncthreads();
        }
        TC_TRACE(27);
        // ---- A operand of layer 0: gathered rows -> TF32 hi/lo, canonical K-major layout (4 loads in flight per thread).
        //      Item i = (chunk j = i / R, row r = i % R): consecutive lanes take consecutive rows of the same 16-byte chunk, so
        //      a quarter-warp's 16-byte stores cover one whole core-matrix column = 128 contiguous bytes, and no division.
        {
            const int K0 = tc.L[0].K_pad, chunks = K0 / 4, total = R * chunks;
            const uint32_t sbo = umma_sbo(K0);
            for (int i0 = tid; i0 < total; i0 += 4 * kTcThreads) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * kTcThreads;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < total) {
                        const int r = i & (R - 1), j = i >> lgR;
                        const float *rp = direct ? ((base + r < a.n) ? a.obs + (size_t)(base + r) * tc.in_dim : nullptr) : rows[r];
                        if (rp && 4 * j < tc.in_dim) v[u] = __ldg(reinterpret_cast<const float4 *>(rp) + j);
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
                        if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + r, 4 * j, sbo)) = l;
                        else *reinterpret_cast<float4 *>(Alo + off) = l;
                    }
                }
            }
        }
        TC_TRACE(2);
        if (!waited) {
            pdl_wait();
            pdl_trigger();
            waited = true;
            if (tid == kCtl && !early_w) stage_image();
        }
        if (!wready) {                                           // first tile: the control warp's barriers and TMEM base become visible
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            tmem = tmem_base_s;
            mbar_wait(&wbar, 0);
            wready = true;
        }
        TC_TRACE(3);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        TC_TRACE(4);

        for (int l = 0; l < tc.n_layers; ++l) {
            const TcLayer T = tc.L[l];
            const uint32_t sbo = umma_sbo(T.K_pad);
            const uint32_t dcol = (uint32_t)(l & 1) * (uint32_t)tc.dstride;
            const uint32_t second = tc.concat ? (uint32_t)T.N_pad : 0u;       // column offset of the hi*lo block
            if (tid == 0) {
                // one thread issues the layer's MMAs.  Descriptors differ only in the 14-bit start-address field:
                // the next K step (two 16-byte chunks = 2*LBO bytes) is +16 in units of 16 B.
                if (stack) issue_3xtf32_stacked(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(W + T.hi_off), sbo), kTcTile, T.N_pad, T.K_pad / 8);
                else issue_3xtf32(tmem + dcol, umma_desc(smem_u32(Ahi), sbo), umma_desc(smem_u32(Alo), sbo),
                                  umma_desc(smem_u32(W + T.hi_off), sbo), umma_desc(smem_u32(W + T.lo_off), sbo), kTcTile, T.N_pad,
                                  T.K_pad / 8, tc.concat != 0);
                umma_commit(&mbar);
            }
            TC_TRACE(5 + 3 * l);
            mbar_wait(&mbar, mphase);
            mphase ^= 1;
            tc_fence_after();
            if (!w2ready) { if (w_split < (uint32_t)tc.img_bytes) mbar_wait(&wbar2, 0); w2ready = true; }   // biases + the next layers' weights
            TC_TRACE(6 + 3 * l);
            const float *bias = bias_all + T.bias_off;
            const int row = quad * 32 + lane;
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + dcol;
            const bool live = quad * 32 < R;                       // this warp's TMEM quadrant holds real rows
            float vpre[32];                                      // this warp's accumulator chunk, loaded while the lo warps park theirs
            if (stack) {                                           // the lo*hi block (rows [R, 2R)) -> scratch -> the hi warps
                if (quad * 32 >= R && quad * 32 < 2 * R)
                    for (int c0 = half * 32; c0 < T.N_pad; c0 += 64) stack_park_lo(taddr, c0, s_lo, row - R);
                if (live && half * 32 < T.N_pad) tmem_ld32_sum(taddr + (uint32_t)(half * 32), second, vpre);
                if (l == 0) TC_TRACE(22);
                __syncthreads();
                if (l == 0) TC_TRACE(23);
            }
            if (l + 1 < tc.n_layers) {
                // hidden layer epilogue: bias + ReLU, re-split, write the next A operand (K_next = N_pad)
                const uint32_t sbon = umma_sbo(T.N_pad);
                for (int c0 = half * 32; live && c0 < T.N_pad; c0 += 64) {
                    float v[32];
                    if (stack) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = vpre[j];
                        stack_add_lo(v, s_lo, row, c0);
                    } else tmem_ld32_sum(taddr + (uint32_t)c0, second, v);
                    if (l == 0) TC_TRACE(24);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 h, lo4;
                        float x0 = fmaxf(v[4 * j + 0] + bias[c0 + 4 * j + 0], 0.f), x1 = fmaxf(v[4 * j + 1] + bias[c0 + 4 * j + 1], 0.f);
                        float x2 = fmaxf(v[4 * j + 2] + bias[c0 + 4 * j + 2], 0.f), x3 = fmaxf(v[4 * j + 3] + bias[c0 + 4 * j + 3], 0.f);
                        tf32_split(x0, h.x, lo4.x); tf32_split(x1, h.y, lo4.y); tf32_split(x2, h.z, lo4.z); tf32_split(x3, h.w, lo4.w);
                        const uint32_t off = umma_off(row, c0 + 4 * j, sbon);
                        *reinterpret_cast<float4 *>(Ahi + off) = h;
                        if (stack) *reinterpret_cast<float4 *>(Ahi + umma_off(R + row, c0 + 4 * j, sbon)) = lo4;
                        else *reinterpret_cast<float4 *>(Alo + off) = lo4;
                    }
                }
                if (l == 0) TC_TRACE(25);
                fence_proxy_async();
                if (l == 0) TC_TRACE(26);
                tc_fence_before();
                __syncthreads();
                tc_fence_after();
                TC_TRACE(7 + 3 * l);
            } else {
                // head epilogue: Q row of this sample, then the mode's output
                if (half == 0 && live) {
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
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j < nA) s += q[j];
                        const float mean = s / (float)nA;
                        float V = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (j == nA) V = q[j];
#pragma unroll
                        for (int j = 0; j < 32; ++j) q[j] = V + q[j] - mean;
                    }
                    int best = 0; float bv = q[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) if (j < nA && q[j] > bv) { bv = q[j]; best = j; }
                    const int b = base + row;
                    if (b < a.n) {
                        if (a.q_out) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (j < nA) a.q_out[(size_t)b * nA + j] = q[j];
                        }
                        if (ACT) {
                            float u; int ra;
                            if (a.u_tape) { u = a.u_tape[b]; ra = a.rand_tape ? a.rand_tape[b] : 0; }
                            else {
                                uint32_t rr[4];
                                Philox::gen(a.key, a.call, (uint64_t)b, rr);
                                u = Philox::u01(rr[0]);
                                ra = (int)(((uint64_t)rr[1] * (uint64_t)nA) >> 32);
                            }
                            a.actions[b] = (u > a.eps || !a.is_train) ? best : ra;      // DuelingDQN_Trainer.py:89-97
                        } else if (!ACT && a.mode == kTcArgmax) {
                            a.actions[b] = best;                                        // DDQN_Trainer.py:94
                        } else {
                            float nq = bv;                                              // DQN_Trainer.py:109
                            if (a.mode == kTcTdGather) {                                // DDQN_Trainer.py:95
                                const int as = a.actions[b];
                                nq = 0.f;
#pragma unroll
                                for (int j = 0; j < 32; ++j) if (j == as) nq = q[j];
                            }
                            a.y_out[b] = s_rew[row] + (a.gamma * nq * (1.f - s_done[row]));   // :99 / :114 / :171
                        }
                    }
                }
                tc_fence_before();
                __syncthreads();
                tc_fence_after();
            }
        }
        if (FUSE_ENV) {
            // the tile's actions are in global memory (written by this CTA before the barrier above)
            __shared__ EnvSmem<32> env_sm;
            for (int sb = 0; sb < R; sb += 32)
                env_block<true, 32, kTcThreads, false, 8>(ef.d, env_sm, base + sb, tid, UAVRL_ACT_DISCRETE27, a.actions, ef.obs_next,
                                                       ef.reward, ef.done, nullptr, nullptr, nullptr);
            __syncthreads();
        }
    }
    TC_TRACE(20);
    tc_fence_before();
    __syncthreads();
    if (warp == kCtl / 32) { tc_fence_after(); tmem_dealloc(tmem_base_s, (uint32_t)tc.tmem_cols); }
    TC_TRACE(21);
}

typedef void (*ForwardKernel)(TcNet, TcArgs, EnvFuse);
template <bool F, bool A, bool S>
static ForwardKernel pick_fwd_d(bool dueling) { return dueling ? tc_forward_kernel_t<F, A, S, true> : tc_forward_kernel_t<F, A, S, false>; }
template <bool F, bool A>
static ForwardKernel pick_fwd_s(bool stack, bool dueling) { return stack ? pick_fwd_d<F, A, true>(dueling) : pick_fwd_d<F, A, false>(dueling); }
// the fused act + env step variant exists for the act mode only
static ForwardKernel pick_forward_kernel(bool fuse_env, bool act, bool stack, bool dueling)
{
    if (fuse_env) return pick_fwd_s<true, true>(stack, dueling);
    return act ? pick_fwd_s<false, true>(stack, dueling) : pick_fwd_s<false, false>(stack, dueling);
}

int launch_tc_forward(uavrl_learner *l, const TcArgs &a_in, cudaStream_t st, const EnvFuse *fuse)
{
    TcArgs a = a_in;
    a.rows_per_tile = (a.n >= 128 * 148) ? 128 : (a.n >= 64 * 148) ? 64 : 32;
    a.n_tiles = (a.n + a.rows_per_tile - 1) / a.rows_per_tile;
    const int grid = a.n_tiles < 148 ? a.n_tiles : 148;
    static const bool trace_on = getenv("UAVRL_TC_TRACE") != nullptr;
    long long *tr = nullptr;
    if (trace_on) { UAVRL_CUDA(cudaMalloc((void **)&tr, 32 * sizeof(long long))); UAVRL_CUDA(cudaMemset(tr, 0, 32 * sizeof(long long))); a.trace = tr; }
    // PDL chain state (see common.cuh): what this kernel may touch before its griddepcontrol.wait
    const int prev = (l->pdl_chain && g_pdl.load()) ? l->pdl_prev : kPdlNone;
    a.pdl = 0;
    if (prev != kPdlNone) {
        a.pdl = kPdlOn;
        if (a.mode == kTcAct) {
            if (prev == kPdlAdam) a.pdl |= kPdlEarlyRows;          // obs frame written by the env step, weights by Adam
            else if (prev == kPdlEnv) a.pdl |= kPdlEarlyWeights;   // collection-only loop: weights untouched, obs just written
        } else if (prev == kPdlEnv || prev == kPdlTd) {
            a.pdl |= kPdlEarlyWeights;                             // neither the env step nor a TD pass writes weight images
        }
    }
    EnvFuse ef;
    memset(&ef, 0, sizeof(ef));
    const bool stack = a.rows_per_tile <= 64 && l->tc.concat != 0 && l->tc.dstride <= 128;
    if (fuse) {
        ef = *fuse;
        UAVRL_CUDA(launch_kernel(pick_forward_kernel(true, a.mode == kTcAct, stack, l->tc.dueling != 0), dim3(grid), dim3(kTcThreads), tc_smem_bytes(l->tc), st, a.pdl != 0, l->tc, a, ef));
        l->pdl_prev = l->pdl_chain ? kPdlEnv : kPdlNone;
    } else {
        UAVRL_CUDA(launch_kernel(pick_forward_kernel(false, a.mode == kTcAct, stack, l->tc.dueling != 0), dim3(grid), dim3(kTcThreads), tc_smem_bytes(l->tc), st, a.pdl != 0, l->tc, a, ef));
        l->pdl_prev = l->pdl_chain ? (a.mode == kTcAct ? kPdlAct : kPdlTd) : kPdlNone;
    }
    UAVRL_LAUNCHED();
    if (trace_on) {
        long long h[32];
        UAVRL_CUDA(cudaStreamSynchronize(st));
        UAVRL_CUDA(cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(tr);
        fprintf(stderr, "[tc_trace] mode=%d n=%d R=%d grid=%d cycles since start:", a.mode, a.n, a.rows_per_tile, grid);
        for (int i = 1; i < 32; ++i) if (h[i]) fprintf(stderr, " [%d]=%lld", i, h[i] - h[0]);
        fprintf(stderr, "\n");
    }
    return 0;
}

int tc_init(uavrl_learner *l)
{
    std::vector<int32_t> hi, lo, hi2, lo2;
    l->tc_ok = false; l->tc_train_ok = false;
    if (tc_build(l->cfg, l->net, l->tc, hi, lo, hi2, lo2) != 0) return 0;
    if (tc_smem_bytes(l->tc) > 227 * 1024) return 0;
    const size_t P = (size_t)l->net.P;
    const size_t img = (size_t)l->tc.train_img_bytes;
    UAVRL_CUDA(cudaMalloc((void **)&l->tc_img_local, img));
    UAVRL_CUDA(cudaMalloc((void **)&l->tc_img_target, img));
    UAVRL_CUDA(cudaMemset(l->tc_img_local, 0, img));
    UAVRL_CUDA(cudaMemset(l->tc_img_target, 0, img));
    int32_t **maps[] = { &l->tc_hi_map, &l->tc_lo_map, &l->tc_hi2_map, &l->tc_lo2_map };
    std::vector<int32_t> *src[] = { &hi, &lo, &hi2, &lo2 };
    for (int i = 0; i < 4; ++i) {
        UAVRL_CUDA(cudaMalloc((void **)maps[i], P * 4));
        UAVRL_CUDA(cudaMemcpy(*maps[i], src[i]->data(), P * 4, cudaMemcpyHostToDevice));
    }
    UAVRL_CUDA(cudaMalloc((void **)&l->y_buf, (size_t)l->cfg.batch_size * 4));
    UAVRL_CUDA(cudaMalloc((void **)&l->astar_buf, (size_t)l->cfg.batch_size * 4));
    for (int ac = 0; ac < 2; ++ac)
        for (int sk = 0; sk < 2; ++sk)
            for (int du = 0; du < 2; ++du)
                UAVRL_CUDA(cudaFuncSetAttribute(pick_forward_kernel(false, ac != 0, sk != 0, du != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)tc_smem_bytes(l->tc)));
    {   // the fused act+step variant carries the env scratch as static shared memory on top: it must still fit one CTA
        l->fuse_ok = true;
        for (int sk = 0; sk < 2; ++sk)
            for (int du = 0; du < 2; ++du) {
                cudaFuncAttributes fa;
                UAVRL_CUDA(cudaFuncGetAttributes(&fa, pick_forward_kernel(true, true, sk != 0, du != 0)));
                if (fa.sharedSizeBytes + tc_smem_bytes(l->tc) > (size_t)227 * 1024) l->fuse_ok = false;
            }
        if (l->fuse_ok)
            for (int sk = 0; sk < 2; ++sk)
                for (int du = 0; du < 2; ++du)
                    UAVRL_CUDA(cudaFuncSetAttribute(pick_forward_kernel(true, true, sk != 0, du != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)tc_smem_bytes(l->tc)));
    }
    l->y_cap = l->cfg.batch_size;
    l->tc_ok = true;
    return tc_train_init(l);
}

}  // namespace uavrl
