// Flash-style self-attention on tcgen05 with error-compensated bf16 operands.
// Replaces F.multi_head_attention_forward -> SDPA with key_padding_mask
// (s3prl/upstream/wav2vec2/wav2vec2_model.py:1146-1168) and, when bias_table is given, WavLM's gated
// relative-position bias (s3prl/upstream/wavlm/modules.py:511-580).
//
// One CTA = one (batch, head, 128-query tile), 192 threads (kHalves = 1):
//   warps 0..3 : softmax warps, thread i owns query row i (TMEM lane i)
//   warp 4     : S warp  (one elected lane): Q/K TMA loads, S_j = Q K_j^T
//   warp 5     : PV warp (one elected lane): V TMA loads, O += P_j V_j
//   (one control warp doing both serialised ~2400 cycles of barrier waits + MMA issue per block)
// Two CTAs are co-resident per SM (<= 168 registers/thread, ~101 KB smem, 256 TMEM columns each).
// TMEM columns: S0 [0,64) | S1 [64,128) | O [128,192) | P_hi [192,224) | P_lo [224,256).
// Per 64-key block j:
//   S_j = Qhi*Khi^T + Qhi*Klo^T + Qlo*Khi^T     tcgen05.mma M=128 N=64 K=64, as soon as its S buffer is free
//   softmax warps: scores (log2 domain: q was scaled by log2(e)/8) -> running max m, p = 2^(s-m), split to
//     bf16 hi/lo and stored to TENSOR MEMORY (tcgen05.st); if any row's max moved, O (which lives in TMEM) is
//     rescaled by alpha with tcgen05.ld / tcgen05.st — only after PV_{j-1} has retired, which it has long before
//   O += Phi*Vhi + Phi*Vlo + Plo*Vhi            tcgen05.mma with A from TMEM, accumulating across blocks
// The softmax warps never wait for PV_j inside the loop, so the tensor pipe, the MUFU pipe and TMA overlap
// (timeline: tools/attn_trace.py). K^T/V blocks: double-buffered, K two blocks ahead (TMA latency ~2000 cycles).
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace s3b {

static constexpr int kQTile = 128;
static constexpr int kKBlk = 64;
static constexpr int kHd = 64;
static constexpr int kQBytes = kQTile * kHd * 2;  // 16 KB per plane
static constexpr int kKBytes = kKBlk * kHd * 2;   // 8 KB per plane
// smem map (1024-aligned): Qhi Qlo | 2 x {Khi Klo} | 2 x {Vhi Vlo} | barriers
static constexpr int kOffQ = 0;
static constexpr int kOffK = 2 * kQBytes;
static constexpr int kStage = 2 * kKBytes;  // one K (or V^T) block, hi + lo planes
static constexpr int kOffV = kOffK + 2 * kStage;
static constexpr int kOffBar = kOffV + 2 * kStage;
// A query row can be shared by kHalves softmax threads (warps w and w + 4 own the same TMEM lane quadrant): thread
// (row, half) handles 64 / kHalves keys of every block and 64 / kHalves columns of O; the row maximum is exchanged
// once per block through shared memory (named barrier per lane quadrant), the row sums stay per-thread partials
// until the end. Measured (same-box A/B, both variants pass the parity tests): 2 threads per row = 1 thread per
// row within 0.5 % — the ~1800-cycle period per 64-key block is set by the two co-resident CTAs' MMAs (12 S MMAs
// that re-read Q from shared memory at 192 B/clk + 12 PV MMAs ~ 960 cycles per block and CTA), not by the softmax
// instruction stream — so the default stays 1 (192 threads).
#ifndef S3B_ATTN_HALVES
#define S3B_ATTN_HALVES 1
#endif
static constexpr int kHalves = S3B_ATTN_HALVES;
static constexpr int kSmWarps = 4 * kHalves;
static constexpr int kCols = kKBlk / kHalves;  // score columns per softmax thread
static constexpr int kDCols = kHd / kHalves;   // O columns per softmax thread
static constexpr int kOffX = kOffBar + 128;    // float [2][2][128] row-max exchange
static constexpr int kAttnSmem = kOffX + 2048 + 1024;  // ~101.5 KB; two CTAs per SM (TMEM: 2 x 256 columns)
static constexpr int kTmemCols = 256;
static constexpr uint32_t kColO = 128, kColPhi = 192, kColPlo = 224;
static constexpr int kAttnThreads = 32 * (kSmWarps + 2);

__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[32]) { tmem_st_32x32(taddr, v); }
__device__ __forceinline__ void tmem_st_cols(uint32_t taddr, const uint32_t (&v)[16]) { tmem_st_32x16(taddr, v); }

template <bool kBias>
__global__ void __launch_bounds__(kAttnThreads, 2) attention_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* bar_k = bar_q + 1;   // [2]  K_j landed in K stage j&1
    uint64_t* bar_v = bar_q + 3;   // [2]  V_j landed in V stage j&1
    uint64_t* bar_s = bar_q + 5;   // [2]  S_j in TMEM (tcgen05.commit)
    uint64_t* bar_pv = bar_q + 7;  //      O += P_j V_j retired (tcgen05.commit)
    uint64_t* bar_p = bar_q + 8;   //      P_j in TMEM, O rescaled, S_j consumed (one arrive per softmax warp)
    uint64_t* bar_sfree = bar_q + 9;  // [2] S buffer j&1 has been read into registers by every row (4 warp arrives)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 11);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int q_tiles = (p.T + kQTile - 1) / kQTile;
    const int n_items = p.B * p.H * q_tiles;  // work items (b, h, 128-query tile); this CTA takes blockIdx.x, + gridDim.x, ...

    if (tid == 0) {
        mbar_init(bar_q, 1);
        mbar_init(&bar_k[0], 1);
        mbar_init(&bar_k[1], 1);
        mbar_init(&bar_v[0], 1);
        mbar_init(&bar_v[1], 1);
        mbar_init(&bar_s[0], 1);
        mbar_init(&bar_s[1], 1);
        mbar_init(bar_pv, 1);
        mbar_init(bar_p, kSmWarps);
        mbar_init(&bar_sfree[0], kSmWarps);
        mbar_init(&bar_sfree[1], kSmWarps);
        fence_mbar_init();
    }
    if (warp == kSmWarps) {
        if (lane == 0) {
            tma_prefetch_desc(&p.q_hi);
            tma_prefetch_desc(&p.k_hi);
            tma_prefetch_desc(&p.vt_hi);
        }
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + kColO;
    pdl_wait();  // prologue overlapped the QKV GEMM's tail; global memory is touched only below
    pdl_launch_dependents();
    const bool tracing = p.trace != nullptr && (int)blockIdx.x == p.trace_block;
#define S3B_TR(role, j, slot)                                                                        \
    do {                                                                                             \
        if (tracing && first_item && (j) < 16) p.trace[((role)*16 + (j)) * 8 + (slot)] = clock64(); \
    } while (0)

    // Persistent over work items: the three roles walk the same (item, key block) sequence; `g` numbers the key blocks
    // of this CTA across items and selects ring stages / barrier phases, so K/V prefetch, the S double buffer and the
    // P / O hand-offs run straight through item boundaries: the next item's Q and first K/V blocks are in flight while
    // the softmax warps finish the current item, and TMEM allocation / barrier init happen once per CTA instead of
    // once per item (a one-item-per-CTA launch spent ~25 % of each CTA in that prologue).
    if (warp == kSmWarps) {
        // ===================== S warp: Q/K loads + S_j = Q K_j^T =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kQTile, 64);
            const uint32_t qa = smem_u32(smem + kOffQ);
            const uint64_t dq_hi = make_smem_desc_sw128(qa), dq_lo = make_smem_desc_sw128(qa + kQBytes);
            int g = 0, it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const bool first_item = it == 0;
                const int bh = item / q_tiles;
                const int q0 = (item - bh * q_tiles) * kQTile;
                const int kv_len = p.kv_len[bh / p.H];
                const int nblk = (kv_len + kKBlk - 1) / kKBlk;
                auto load_k = [&](int j) {  // K_j of this item -> K stage (g+j)&1
                    const int st_i = (g + j) & 1;
                    uint8_t* st = smem + kOffK + st_i * kStage;
                    mbar_arrive_expect_tx(&bar_k[st_i], 2 * kKBytes);
                    tma_load_3d(st, &p.k_hi, &bar_k[st_i], 0, j * kKBlk, bh);
                    tma_load_3d(st + kKBytes, &p.k_lo, &bar_k[st_i], 0, j * kKBlk, bh);
                };
                // Q smem and both K stages are free: the previous item's S MMAs have all retired (end of its loop)
                mbar_arrive_expect_tx(bar_q, 2 * kQBytes);
                tma_load_3d(smem + kOffQ, &p.q_hi, bar_q, 0, q0, bh);
                tma_load_3d(smem + kOffQ + kQBytes, &p.q_lo, bar_q, 0, q0, bh);
                load_k(0);
                if (nblk > 1) load_k(1);
                mbar_wait(bar_q, (uint32_t)(it & 1));
                for (int j = 0; j < nblk; ++j) {
                    const int gg = g + j;
                    S3B_TR(0, j, 0);
                    mbar_wait(&bar_k[gg & 1], (uint32_t)((gg >> 1) & 1));
                    // S buffer gg&1 is free once every row has read S_{gg-2} into registers
                    if (gg >= 2) mbar_wait(&bar_sfree[gg & 1], (uint32_t)(((gg >> 1) - 1) & 1));
                    tc_fence_after();
                    S3B_TR(0, j, 1);
                    const uint32_t ka = smem_u32(smem + kOffK + (gg & 1) * kStage);
                    const uint64_t dk_hi = make_smem_desc_sw128(ka), dk_lo = make_smem_desc_sw128(ka + kKBytes);
                    const uint32_t d = tmem_base + (uint32_t)(gg & 1) * 64u;
#pragma unroll
                    for (int k = 0; k < kHd / 16; ++k) {
                        const uint64_t ko = (uint64_t)(2 * k);
                        umma_bf16(d, dq_lo + ko, dk_hi + ko, idesc, k != 0 ? 1u : 0u);
                        umma_bf16(d, dq_hi + ko, dk_lo + ko, idesc, 1u);
                        umma_bf16(d, dq_hi + ko, dk_hi + ko, idesc, 1u);
                    }
                    umma_commit(&bar_s[gg & 1]);
                    S3B_TR(0, j, 2);
                    if (j + 2 < nblk) {  // K_{j+2} replaces K_j as soon as S_j has retired
                        mbar_wait(&bar_s[gg & 1], (uint32_t)((gg >> 1) & 1));
                        load_k(j + 2);
                    }
                    S3B_TR(0, j, 3);
                }
                // the last (up to two) S MMAs of the item were not waited for above: Q and the K stages are reused next
                for (int j = (nblk >= 2 ? nblk - 2 : 0); j < nblk; ++j) {
                    const int gg = g + j;
                    mbar_wait(&bar_s[gg & 1], (uint32_t)((gg >> 1) & 1));
                }
                g += nblk;
            }
        }
    } else if (warp == kSmWarps + 1) {
        // ===================== PV warp: V loads + O += P_j V_j =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kQTile, 64);
            int g = 0, it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const bool first_item = it == 0;
                const int bh = item / q_tiles;
                const int kv_len = p.kv_len[bh / p.H];
                const int nblk = (kv_len + kKBlk - 1) / kKBlk;
                auto load_v = [&](int j) {  // V^T_j of this item -> V stage (g+j)&1
                    const int st_i = (g + j) & 1;
                    uint8_t* st = smem + kOffV + st_i * kStage;
                    mbar_arrive_expect_tx(&bar_v[st_i], 2 * kKBytes);
                    tma_load_3d(st, &p.vt_hi, &bar_v[st_i], j * kKBlk, 0, bh);
                    tma_load_3d(st + kKBytes, &p.vt_lo, &bar_v[st_i], j * kKBlk, 0, bh);
                };
                load_v(0);  // both V stages are free: the previous item's last PV has retired (end of its loop)
                if (nblk > 1) load_v(1);
                for (int j = 0; j < nblk; ++j) {
                    const int gg = g + j;
                    mbar_wait(&bar_v[gg & 1], (uint32_t)((gg >> 1) & 1));  // V_j landed long ago
                    S3B_TR(0, j, 4);
                    mbar_wait(bar_p, (uint32_t)(gg & 1));  // P_j in TMEM, O rescaled, S_j consumed by every row
                    tc_fence_after();
                    S3B_TR(0, j, 5);
                    const uint32_t va = smem_u32(smem + kOffV + (gg & 1) * kStage);
                    const uint64_t dv_hi = make_smem_desc_sw128(va), dv_lo = make_smem_desc_sw128(va + kKBytes);
#pragma unroll
                    for (int k = 0; k < kKBlk / 16; ++k) {
                        const uint64_t ko = (uint64_t)(2 * k);
                        // 16 keys per MMA = 8 TMEM columns of packed bf16 pairs
                        const uint32_t a_hi = tmem_base + kColPhi + 8u * k, a_lo = tmem_base + kColPlo + 8u * k;
                        umma_bf16_ts(tmem_o, a_lo, dv_hi + ko, idesc, (j | k) != 0 ? 1u : 0u);
                        umma_bf16_ts(tmem_o, a_hi, dv_lo + ko, idesc, 1u);
                        umma_bf16_ts(tmem_o, a_hi, dv_hi + ko, idesc, 1u);
                    }
                    umma_commit(bar_pv);
                    S3B_TR(0, j, 6);
                    if (j + 2 < nblk || j == nblk - 1) {
                        // V_{j+2} replaces V_j once PV_j has retired; after the item's last block the wait frees both
                        // stages for the next item (PV_{nblk-2} retired before it: same accumulator, issue order)
                        mbar_wait(bar_pv, (uint32_t)(gg & 1));
                        if (j + 2 < nblk) load_v(j + 2);
                    }
                }
                g += nblk;
            }
        }
    } else {
        // ===================== softmax warps =====================
        const int quad = warp & 3;   // TMEM lane quadrant
        const int half = warp >> 2;  // which key / O-column share of the row
        const int row_in_tile = quad * 32 + lane;
        const uint32_t lane_off = ((uint32_t)(quad * 32)) << 16;
        float* xch = reinterpret_cast<float*>(smem + kOffX);
        int g = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const bool first_item = it == 0;
            const int bh = item / q_tiles;
            const int q0 = (item - bh * q_tiles) * kQTile;
            const int b = bh / p.H;
            const int h = bh - b * p.H;
            const int kv_len = p.kv_len[b];
            const int nblk = (kv_len + kKBlk - 1) / kKBlk;
            const int q_row = q0 + row_in_tile;
            const bool row_ok = q_row < p.T;
            float gate = 0.f;
            const float* brow = nullptr;
            if (kBias) {
                gate = (p.gate == nullptr) ? 1.0f : (row_ok ? p.gate[((size_t)b * p.H + h) * p.T + q_row] : 0.f);
                brow = p.bias_table + (size_t)h * p.bias_stride + (p.bias_center - (row_ok ? q_row : 0));  // index by key k
            }
            float m_run = -INFINITY;  // running row max (log2 domain)
            float l_run = 0.f;        // this thread's share of the row sum

            for (int j = 0; j < nblk; ++j) {
                const int gg = g + j;
                mbar_wait(&bar_s[gg & 1], (uint32_t)((gg >> 1) & 1));
                __syncwarp();
                tc_fence_after();
                if (tid == 0) S3B_TR(1, j, 0);
                float s[kCols];
                {
                    const uint32_t ts = tmem_base + (uint32_t)(gg & 1) * 64u + lane_off + (uint32_t)(half * kCols);
#pragma unroll
                    for (int c = 0; c < kCols; c += 32) {
                        uint32_t v0[32];
                        tmem_ld_32x32(ts + (uint32_t)c, v0);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) s[c + i] = __uint_as_float(v0[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_sfree[gg & 1]);  // the S warp may overwrite this buffer with S_{gg+2}
                const int kbase = j * kKBlk + half * kCols;
                if (kBias) {
                    if (kbase + kCols <= p.T) {  // block-uniform: every key of the block has a table entry
#pragma unroll
                        for (int i = 0; i < kCols; ++i) s[i] = fmaf(gate, __ldg(brow + kbase + i), s[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < kCols; ++i) {
                            const int kk = kbase + i;
                            s[i] = fmaf(gate, (kk < p.T) ? __ldg(brow + kk) : 0.f, s[i]);
                        }
                    }
                }
                if (j * kKBlk + kKBlk > kv_len) {  // block-uniform: only the last block is partially masked
#pragma unroll
                    for (int i = 0; i < kCols; ++i)
                        if (kbase + i >= kv_len) s[i] = -INFINITY;
                }
                float mx = fmaxf(s[0], s[1]);
#pragma unroll
                for (int i = 2; i < kCols; i += 2) mx = fmax3(mx, s[i], s[i + 1]);
                if constexpr (kHalves == 2) {
                    // double-buffered by block parity: the partner reads buffer gg&1 before it reaches the barrier of
                    // block gg+1, and this thread rewrites it only in block gg+2
                    xch[((gg & 1) * 2 + half) * 128 + row_in_tile] = mx;
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
                    mx = fmaxf(mx, xch[((gg & 1) * 2 + (half ^ 1)) * 128 + row_in_tile]);
                }
                const float m_new = fmaxf(m_run, mx);  // finite: key j*64 is always valid
                const float alpha = fast_exp2(m_run - m_new);
                float psum0 = 0.f, psum1 = 0.f;
                uint32_t hw[kCols / 2], lw[kCols / 2];
#pragma unroll
                for (int e = 0; e < kCols / 2; ++e) {  // 10 instructions per key pair: FMNMX3 above, FADD2, 2 MUFU, FADD2, split (5)
                    float d0, d1;
                    fsub2(d0, d1, s[2 * e], s[2 * e + 1], m_new, m_new);
                    const float p0 = fast_exp2(d0);
                    const float p1 = fast_exp2(d1);
                    fadd2(psum0, psum1, psum0, psum1, p0, p1);
                    split_pack2(p0, p1, hw[e], lw[e]);
                }
                l_run = fmaf(l_run, alpha, psum0 + psum1);
                m_run = m_new;
                if (tid == 0) S3B_TR(1, j, 1);

                if (j > 0) {
                    // P_{j-1} and O are still owned by PV_{j-1} until it retires (issued ~one softmax phase ago)
                    mbar_wait(bar_pv, (uint32_t)((gg - 1) & 1));
                    __syncwarp();
                    tc_fence_after();
                    if (tid == 0) S3B_TR(1, j, 2);
                    if (!__all_sync(0xffffffffu, alpha == 1.0f)) {  // some row's max moved: rescale O in place
#pragma unroll
                        for (int c = 0; c < kDCols; c += 32) {
                            uint32_t v[32];
                            const uint32_t ta = tmem_o + lane_off + (uint32_t)(half * kDCols + c);
                            tmem_ld_32x32(ta, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                            tmem_st_32x32(ta, v);
                        }
                    }
                }
                // (j == 0: the previous item's last PV was waited for before its O was read, below)
                // packed bf16 pairs: key 2c, 2c+1 of the block in column c of the P_hi / P_lo planes
                tmem_st_cols(tmem_base + lane_off + kColPhi + (uint32_t)(half * (kCols / 2)), hw);
                tmem_st_cols(tmem_base + lane_off + kColPlo + (uint32_t)(half * (kCols / 2)), lw);
                tmem_st_wait();     // P_j and the rescaled O are in tensor memory
                tc_fence_before();  // orders this thread's tcgen05.ld / tcgen05.st before the hand-off
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p);
                if (tid == 0) S3B_TR(1, j, 3);
            }

            // O is complete once the item's last PV has retired
            const int g_last = g + nblk - 1;
            mbar_wait(bar_pv, (uint32_t)(g_last & 1));
            __syncwarp();
            tc_fence_after();
            float l_tot = l_run;
            if constexpr (kHalves == 2) {
                // buffer (g_last+1)&1 was last read in block g_last-1, which the partner left before the barrier of g_last
                xch[(((g_last + 1) & 1) * 2 + half) * 128 + row_in_tile] = l_run;
                asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
                l_tot += xch[(((g_last + 1) & 1) * 2 + (half ^ 1)) * 128 + row_in_tile];
            }
            const float inv = 1.0f / l_tot;
            const size_t off = ((size_t)b * p.T + (row_ok ? q_row : 0)) * (size_t)p.D + (size_t)h * kHd + half * kDCols;
#pragma unroll
            for (int c = 0; c < kDCols; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(half * kDCols + c), v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        float y[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(v[i + e]) * inv;
                        store_planes8(p.ctx, y, off + c + i);
                    }
                }
            }
            // the next item's first PV (accumulate = 0) overwrites O: it is gated by this warp's arrive on bar_p for
            // that item's block 0, which comes after the tcgen05.ld above (tc_fence_before before the arrive)
            g += nblk;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kSmWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

cudaError_t launch_attention(const AttnParams& p, cudaStream_t s) {
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.current();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kAttnSmem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int q_tiles = (p.T + kQTile - 1) / kQTile;
    const int n_items = p.B * p.H * q_tiles;
    if (n_items <= 0) return cudaSuccess;
    // persistent: two co-resident CTAs per SM walk the items (S3B_ATTN_PERSIST=0: one CTA per item, the round-1 launch)
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (sm_count[dev & 63] == 0) cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    static int persist = -1;
    if (persist < 0) {
        const char* e = getenv("S3B_ATTN_PERSIST");
        persist = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int slots = 2 * (sm_count[dev & 63] > 0 ? sm_count[dev & 63] : 148);
    const int grid = (persist && n_items > slots) ? slots : n_items;
    if (p.bias_table != nullptr)
        return launch_pdl(attention_kernel<true>, dim3(grid), dim3(kAttnThreads), kAttnSmem, s, p);
    return launch_pdl(attention_kernel<false>, dim3(grid), dim3(kAttnThreads), kAttnSmem, s, p);
}

}  // namespace s3b
