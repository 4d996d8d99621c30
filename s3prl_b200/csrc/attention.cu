// Flash-style self-attention on tcgen05 with error-compensated bf16 operands.
// Replaces F.multi_head_attention_forward -> SDPA with key_padding_mask
// (s3prl/upstream/wav2vec2/wav2vec2_model.py:1146-1168) and, when bias_table/gate are given, WavLM's
// gated relative-position bias (s3prl/upstream/wavlm/modules.py:511-580).
//
// One CTA = one (batch, head, 128-query tile); thread i owns query row i (TMEM lane i).
// Per 64-key block:
//   S = Qhi*Khi^T + Qhi*Klo^T + Qlo*Khi^T       tcgen05.mma  M=128 N=64 K=64   (TMEM cols [0,64))
//   online softmax in fp32 registers (row per thread), P split to bf16 hi/lo -> 128B-swizzled smem
//   PV = Phi*Vhi + Phi*Vlo + Plo*Vhi            tcgen05.mma  M=128 N=64 K=64   (TMEM cols [64,128))
//   O = O*alpha + PV in registers
// K / V^T blocks are double-buffered through TMA; q was pre-scaled by 1/sqrt(64) in the QKV epilogue.
#include <math.h>

#include "common.cuh"
#include "kernels.cuh"

namespace s3b {

static constexpr int kQTile = 128;
static constexpr int kKBlk = 64;
static constexpr int kHd = 64;
static constexpr int kQBytes = kQTile * kHd * 2;   // 16 KB per hi / lo
static constexpr int kKBytes = kKBlk * kHd * 2;    // 8 KB
static constexpr int kPBytes = kQTile * kKBlk * 2; // 16 KB
// smem map (1024-aligned): Qhi Qlo | 2 x {Khi Klo Vhi Vlo} | Phi Plo | barriers
static constexpr int kOffQ = 0;
static constexpr int kOffKV = 2 * kQBytes;
static constexpr int kKVStage = 4 * kKBytes;
static constexpr int kOffP = kOffKV + 2 * kKVStage;
static constexpr int kOffBar = kOffP + 2 * kPBytes;
static constexpr int kAttnSmem = kOffBar + 128 + 1024;
static constexpr float kLog2e = 1.4426950408889634f;

__global__ void __launch_bounds__(128, 1) attention_kernel(const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bar_q = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* bar_kv = bar_q + 1;  // [2]
    uint64_t* bar_s = bar_q + 3;
    uint64_t* bar_pv = bar_q + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 5);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int q_tiles = (p.T + kQTile - 1) / kQTile;
    const int bh = blockIdx.x / q_tiles;
    const int q0 = (blockIdx.x - bh * q_tiles) * kQTile;
    const int b = bh / p.H;
    const int h = bh - b * p.H;
    const int kv_len = p.kv_len[b];
    const int nblk = (kv_len + kKBlk - 1) / kKBlk;

    if (tid == 0) {
        tma_prefetch_desc(&p.q_hi);
        tma_prefetch_desc(&p.k_hi);
        tma_prefetch_desc(&p.vt_hi);
        mbar_init(bar_q, 1);
        mbar_init(&bar_kv[0], 1);
        mbar_init(&bar_kv[1], 1);
        mbar_init(bar_s, 1);
        mbar_init(bar_pv, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 64;
    const uint32_t lane_off = ((uint32_t)(warp * 32)) << 16;

    auto load_kv = [&](int j) {
        uint8_t* st = smem + kOffKV + (j & 1) * kKVStage;
        uint64_t* bar = &bar_kv[j & 1];
        mbar_arrive_expect_tx(bar, 4 * kKBytes);
        tma_load_3d(st, &p.k_hi, bar, 0, j * kKBlk, bh);
        tma_load_3d(st + kKBytes, &p.k_lo, bar, 0, j * kKBlk, bh);
        tma_load_3d(st + 2 * kKBytes, &p.vt_hi, bar, j * kKBlk, 0, bh);
        tma_load_3d(st + 3 * kKBytes, &p.vt_lo, bar, j * kKBlk, 0, bh);
    };

    if (tid == 0) {
        mbar_arrive_expect_tx(bar_q, 2 * kQBytes);
        tma_load_3d(smem + kOffQ, &p.q_hi, bar_q, 0, q0, bh);
        tma_load_3d(smem + kOffQ + kQBytes, &p.q_lo, bar_q, 0, q0, bh);
        load_kv(0);
    }

    const int q_row = q0 + tid;
    const bool row_ok = q_row < p.T;
    const bool has_bias = (p.bias_table != nullptr);
    float gate = 0.f;
    const float* brow = nullptr;
    if (has_bias) {
        gate = (p.gate == nullptr) ? 1.0f : (row_ok ? p.gate[((size_t)b * p.H + h) * p.T + q_row] : 0.f);
        // table index for key k: k - q + T - 1
        brow = p.bias_table + (size_t)h * (2 * p.T - 1) + (p.T - 1 - (row_ok ? q_row : 0));
    }

    float o[kHd];
#pragma unroll
    for (int d = 0; d < kHd; ++d) o[d] = 0.f;
    float m_run = -INFINITY;  // running max, in log2 domain (s * log2e)
    float l_run = 0.f;

    const uint32_t idesc = make_idesc_bf16(kQTile, 64);
    uint8_t* p_hi_s = smem + kOffP;
    uint8_t* p_lo_s = smem + kOffP + kPBytes;
    // 128B-swizzle placement of this thread's row inside a [128 x 64] bf16 K-major tile
    const uint32_t row_off = (uint32_t)(tid >> 3) * 1024u + (uint32_t)(tid & 7) * 128u;
    const uint32_t row_xor = (uint32_t)(tid & 7);

    for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        if (tid == 0) {
            if (j + 1 < nblk) load_kv(j + 1);  // buffer (j+1)&1 was released when PV_{j-1} completed
            if (j == 0) mbar_wait(bar_q, 0);
            mbar_wait(&bar_kv[buf], (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            const uint32_t qa = smem_u32(smem + kOffQ);
            const uint32_t kb = smem_u32(smem + kOffKV + buf * kKVStage);
            const uint64_t dq_hi = make_smem_desc_sw128(qa), dq_lo = make_smem_desc_sw128(qa + kQBytes);
            const uint64_t dk_hi = make_smem_desc_sw128(kb), dk_lo = make_smem_desc_sw128(kb + kKBytes);
#pragma unroll
            for (int k = 0; k < kHd / 16; ++k) {
                const uint64_t ko = (uint64_t)(2 * k);
                umma_bf16(tmem_s, dq_lo + ko, dk_hi + ko, idesc, k != 0 ? 1u : 0u);
                umma_bf16(tmem_s, dq_hi + ko, dk_lo + ko, idesc, 1u);
                umma_bf16(tmem_s, dq_hi + ko, dk_hi + ko, idesc, 1u);
            }
            umma_commit(bar_s);
        }
        mbar_wait(bar_s, (uint32_t)(j & 1));
        __syncwarp();
        tc_fence_after();

        // ---- online softmax on this thread's row -------------------------------------------------
        float s[kKBlk];
        {
            uint32_t v[32];
            tmem_ld_32x32(tmem_s + lane_off, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(v[i]);
            tmem_ld_32x32(tmem_s + lane_off + 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) s[32 + i] = __uint_as_float(v[i]);
        }
        const int kbase = j * kKBlk;
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < kKBlk; ++i) {
            float x = s[i];
            if (has_bias) {
                const int kk = kbase + i;
                x = fmaf(gate, (kk < p.T) ? __ldg(brow + kk) : 0.f, x);
            }
            x = (kbase + i < kv_len) ? x * kLog2e : -INFINITY;
            s[i] = x;
            mx = fmaxf(mx, x);
        }
        const float m_new = fmaxf(m_run, mx);  // finite: key 0 of block 0 is always valid (kv_len >= 1)
        const float alpha = exp2f(m_run - m_new);
        float psum = 0.f;
#pragma unroll
        for (int i = 0; i < kKBlk; i += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p0 = exp2f(s[i + 2 * e] - m_new);
                const float p1 = exp2f(s[i + 2 * e + 1] - m_new);
                psum += p0 + p1;
                split_pack2(p0, p1, hw[e], lw[e]);
            }
            const uint32_t chunk = ((uint32_t)(i >> 3) ^ row_xor) * 16u;
            *reinterpret_cast<uint4*>(p_hi_s + row_off + chunk) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(p_lo_s + row_off + chunk) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        l_run = l_run * alpha + psum;
        m_run = m_new;

        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor-core async proxy
        tc_fence_before();
        __syncthreads();

        if (tid == 0) {
            tc_fence_after();
            const uint32_t pa = smem_u32(p_hi_s);
            const uint32_t vb = smem_u32(smem + kOffKV + buf * kKVStage + 2 * kKBytes);
            const uint64_t dp_hi = make_smem_desc_sw128(pa), dp_lo = make_smem_desc_sw128(pa + kPBytes);
            const uint64_t dv_hi = make_smem_desc_sw128(vb), dv_lo = make_smem_desc_sw128(vb + kKBytes);
#pragma unroll
            for (int k = 0; k < kKBlk / 16; ++k) {
                const uint64_t ko = (uint64_t)(2 * k);
                umma_bf16(tmem_o, dp_lo + ko, dv_hi + ko, idesc, k != 0 ? 1u : 0u);
                umma_bf16(tmem_o, dp_hi + ko, dv_lo + ko, idesc, 1u);
                umma_bf16(tmem_o, dp_hi + ko, dv_hi + ko, idesc, 1u);
            }
            umma_commit(bar_pv);
        }
        mbar_wait(bar_pv, (uint32_t)(j & 1));
        __syncwarp();
        tc_fence_after();
        {
            uint32_t v[32];
            tmem_ld_32x32(tmem_o + lane_off, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
            tmem_ld_32x32(tmem_o + lane_off + 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[32 + i] = fmaf(o[32 + i], alpha, __uint_as_float(v[i]));
        }
        tc_fence_before();
        __syncthreads();  // S / PV columns and the P tile may be overwritten by the next iteration
        tc_fence_after();
    }

    if (row_ok) {
        const float inv = 1.0f / l_run;
        const size_t off = ((size_t)b * p.T + q_row) * (size_t)p.D + (size_t)h * kHd;
        uint4* dh = reinterpret_cast<uint4*>(p.ctx_hi + off);
        uint4* dl = reinterpret_cast<uint4*>(p.ctx_lo + off);
#pragma unroll
        for (int i = 0; i < kHd; i += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_pack2(o[i + 2 * e] * inv, o[i + 2 * e + 1] * inv, hw[e], lw[e]);
            dh[i >> 3] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            dl[i >> 3] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

cudaError_t launch_attention(const AttnParams& p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e =
            cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int q_tiles = (p.T + kQTile - 1) / kQTile;
    const int grid = p.B * p.H * q_tiles;
    if (grid <= 0) return cudaSuccess;
    attention_kernel<<<grid, 128, kAttnSmem, s>>>(p);
    return cudaGetLastError();
}

}  // namespace s3b
