// Persistent, warp-specialised tcgen05 GEMM with error-compensated bf16 operands (3 MMAs per product):
//   D[m][n] = sum_k A[m][k] * W[n][k],  A = A_hi + A_lo, W = W_hi + W_lo (bf16),  fp32 accumulation in TMEM
//   D += A_hi*W_hi + A_hi*W_lo + A_lo*W_hi
//
// Replaces the reference's fp32 torch ops (cited per call site in model.cu):
//   nn.Conv1d(512,512,k,2) blocks      s3prl/upstream/wav2vec2/wav2vec2_model.py:2879,2927-2934
//   post_extract_proj                  s3prl/upstream/hubert/hubert_model.py:489-490
//   pos_conv (grouped Conv1d k=128)    s3prl/upstream/wav2vec2/wav2vec2_model.py:2937-2953,3064-3067
//   q/k/v/out projections, fc1, fc2    s3prl/upstream/wav2vec2/wav2vec2_model.py:1146-1168,3260-3322
//
// Roles (256 threads, one CTA per SM, persistent over output tiles):
//   warp 0 lane 0 : TMA producer  (A_hi, A_lo, W_hi, W_lo boxes -> 128B-swizzled smem ring)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma kind::f16, M=128, N=umma_n, K=16 per instruction)
//   warp 2        : TMEM allocator (512 columns = two 128 x 256 fp32 accumulator stages)
//   warps 4..7    : epilogue      (tcgen05.ld -> bias / GELU / residual / mask -> fp32 and/or bf16 hi+lo stores)
#include <stdio.h>

#include "common.cuh"
#include "gemm.cuh"

namespace s3b {

static constexpr int kBlockM = 128;
static constexpr int kAccCols = 256;             // TMEM columns per accumulator stage
static constexpr int kThreads = 256;

// BLOCK_K bf16 elements per pipeline stage = one swizzle row: 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B.
// 128x256 tiles use BLOCK_K = 32 so that four 48 KB stages fit (three loads in flight cover the TMA latency;
// the first version had two 96 KB stages and ran the tensor pipe at ~60 %, profiles/r1a).
template <int BLOCK_N, int BLOCK_K>
struct GemmCfg {
    static constexpr int kBlockK = BLOCK_K;
    static constexpr int kATileBytes = kBlockM * BLOCK_K * 2;
    static constexpr int kBTileBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void store_f32x32(float* dst, const float (&x)[32]) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
}

template <int NC>
__device__ __forceinline__ void store_split(__nv_bfloat16* dhi, __nv_bfloat16* dlo, const float (&x)[NC]) {
    uint32_t h[NC / 2], l[NC / 2];
#pragma unroll
    for (int j = 0; j < NC / 2; ++j) split_pack2(x[2 * j], x[2 * j + 1], h[j], l[j]);
    uint4* h4 = reinterpret_cast<uint4*>(dhi);
    uint4* l4 = reinterpret_cast<uint4*>(dlo);
#pragma unroll
    for (int j = 0; j < NC / 8; ++j) {
        h4[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
        l4[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
    }
}

// Epilogue for NC (16 or 32) consecutive accumulator columns of one output row.
// residual row segment for NC columns, issued early so that its global-load latency overlaps the TMEM load and
// the previous chunk's math (out_proj / fc2 epilogues were stalling on these loads, profiles/r1c)
template <int NC>
__device__ __forceinline__ void load_residual(const GemmParams& p, bool row_ok, size_t m, int col, float4 (&r)[NC / 4]) {
    if (p.residual != nullptr && row_ok && !p.qkv_mode) {
        const float4* r4 = reinterpret_cast<const float4*>(p.residual + m * (size_t)p.ldo + col);
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) r[j] = __ldg(r4 + j);
    }
}

template <int NC>
__device__ __forceinline__ void epilogue_cols(const GemmParams& p, const uint32_t (&v)[NC], bool row_ok, size_t m,
                                              int col, bool masked, const float4 (&res)[NC / 4]) {
    float x[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) x[j] = __uint_as_float(v[j]);
    if (p.bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
            const float4 b = __ldg(b4 + j);
            x[4 * j] += b.x, x[4 * j + 1] += b.y, x[4 * j + 2] += b.z, x[4 * j + 3] += b.w;
        }
    }
    if (p.gelu) {
#pragma unroll
        for (int j = 0; j < NC; ++j) x[j] = gelu_erf(x[j]);
    }
    if (!row_ok) return;

    if (p.qkv_mode) {
        // m = b*T + t ; col = which*D + h*64 + d  (NC-aligned chunks never straddle a head)
        const int b = (int)(m / (size_t)p.T);
        const int t = (int)(m - (size_t)b * p.T);
        const int which = col / p.D;
        const int within = col - which * p.D;
        const int h = within >> 6;
        const int d = within & 63;
        const size_t bh = (size_t)b * p.H + h;
        if (which < 2) {
            if (which == 0) {
#pragma unroll
                for (int j = 0; j < NC; ++j) x[j] *= p.q_scale;
            }
            __nv_bfloat16* dh = (which == 0 ? p.q_hi : p.k_hi) + (bh * p.T + t) * 64 + d;
            __nv_bfloat16* dl = (which == 0 ? p.q_lo : p.k_lo) + (bh * p.T + t) * 64 + d;
            store_split<NC>(dh, dl, x);
        } else {
            __nv_bfloat16* dh = p.vt_hi + (bh * 64 + d) * (size_t)p.Tp + t;
            __nv_bfloat16* dl = p.vt_lo + (bh * 64 + d) * (size_t)p.Tp + t;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                __nv_bfloat16 hi, lo;
                split_bf16(x[j], hi, lo);
                dh[(size_t)j * p.Tp] = hi;
                dl[(size_t)j * p.Tp] = lo;
            }
        }
        return;
    }

    const size_t off = m * (size_t)p.ldo + col;
    if (p.out_pre != nullptr) {
        float4* d4 = reinterpret_cast<float4*>(p.out_pre + off);
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) d4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    }
    if (p.residual != nullptr) {
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) {
            const float4 r = res[j];
            x[4 * j] += r.x, x[4 * j + 1] += r.y, x[4 * j + 2] += r.z, x[4 * j + 3] += r.w;
        }
    }
    if (masked) {
#pragma unroll
        for (int j = 0; j < NC; ++j) x[j] = 0.0f;
    }
    if (p.out_f32 != nullptr) {
        float4* d4 = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
        for (int j = 0; j < NC / 4; ++j) d4[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
    }
    if (p.out_hi != nullptr) store_split<NC>(p.out_hi + off, p.out_lo + off, x);
}


// ------------------------------------------------------------------------------------------------
// Coalesced epilogue of one 32-row x 32-column accumulator chunk (CTA-pair kernel).
// tcgen05.ld hands every thread one ROW (32 consecutive columns): storing that directly makes each warp store
// touch 32 different 128-byte lines (16 B each), and the LSU then needs ~20-29 k cycles per 128 x 256 tile —
// longer than the 12 k-blocks of MMAs of a K=768 GEMM (timeline in profiles/r1e_gemm_trace.txt). The chunk is
// therefore transposed through a per-warp 4 KB shared-memory buffer (XOR-swizzled, conflict-free both ways):
// afterwards lane l owns columns 4*(l&7)..+3 of rows it*4 + (l>>3), it = 0..7, so each warp-level residual load /
// fp32 store covers 4 rows x 128 contiguous bytes (bf16 planes: 4 rows x 64 B).
// ------------------------------------------------------------------------------------------------
struct EpiRows {           // per-tile row bookkeeping of one lane in the transposed layout
    size_t off[8];         // element offset of the output row start (normal: m*ldo, QKV q/k: (b*H*T + t)*64)
    uint32_t ok_bits;      // bit it: row exists
    uint32_t mask_bits;    // bit it: row is a padded frame (zeroed)
};

__device__ __forceinline__ void epi_rows_init(const GemmParams& p, EpiRows& er, int batch, int row_base, int lane,
                                              bool masked_thread_row) {
    er.ok_bits = 0, er.mask_bits = 0;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int src = it * 4 + (lane >> 3);
        const int row = row_base + src;
        const bool ok = row < p.rows_per_batch;
        const size_t m = (size_t)batch * p.out_rows_per_batch + (ok ? row : 0);
        if (p.qkv_mode) {
            const int b = (int)(m / (size_t)p.T);
            const int t = (int)(m - (size_t)b * p.T);
            er.off[it] = ((size_t)b * p.H * p.T + t) * 64;
        } else {
            er.off[it] = m * (size_t)p.ldo;
        }
        er.ok_bits |= (ok ? 1u : 0u) << it;
        er.mask_bits |= (__shfl_sync(0xffffffffu, masked_thread_row ? 1 : 0, src) ? 1u : 0u) << it;
    }
}

__device__ __forceinline__ void epi_load_residual(const GemmParams& p, const EpiRows& er, int col, int lane,
                                                  float4 (&r)[8]) {
    if (p.residual != nullptr && !p.qkv_mode) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
            if ((er.ok_bits >> it) & 1u)
                r[it] = __ldg(reinterpret_cast<const float4*>(p.residual + er.off[it] + col) + (lane & 7));
    }
}

// v: this thread's row (32 columns starting at `col`); stage: this warp's 4 KB buffer
__device__ __forceinline__ void epi_chunk_coalesced(const GemmParams& p, const EpiRows& er, const uint32_t (&v)[32],
                                                    uint8_t* stage, int col, int lane, const float4 (&res)[8],
                                                    const float4& bias) {
    const uint32_t sbase = smem_u32(stage);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t a = sbase + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                     "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                     : "memory");
    }
    __syncwarp();
    const int cs = lane & 7;
    // QKV scatter: 32-column chunks never straddle q/k/v or a head
    int which = 0;
    size_t head_off = 0;
    if (p.qkv_mode) {
        which = col / p.D;
        const int within = col - which * p.D;
        head_off = (size_t)(within >> 6) * p.T * 64 + (within & 63);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const uint32_t a = sbase + (uint32_t)r * 128u + (uint32_t)((cs ^ (r & 7)) << 4);
        float4 x;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(a));
        x.x += bias.x, x.y += bias.y, x.z += bias.z, x.w += bias.w;
        if (p.gelu) x.x = gelu_erf(x.x), x.y = gelu_erf(x.y), x.z = gelu_erf(x.z), x.w = gelu_erf(x.w);
        if (!((er.ok_bits >> it) & 1u)) continue;
        if (p.qkv_mode) {
            if (which == 0) x.x *= p.q_scale, x.y *= p.q_scale, x.z *= p.q_scale, x.w *= p.q_scale;
            uint32_t h0, l0, h1, l1;
            split_pack2(x.x, x.y, h0, l0);
            split_pack2(x.z, x.w, h1, l1);
            const size_t o = er.off[it] + head_off + 4 * cs;
            *reinterpret_cast<uint2*>((which == 0 ? p.q_hi : p.k_hi) + o) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>((which == 0 ? p.q_lo : p.k_lo) + o) = make_uint2(l0, l1);
            continue;
        }
        const size_t o = er.off[it] + col + 4 * cs;
        if (p.out_pre != nullptr) *reinterpret_cast<float4*>(p.out_pre + o) = x;
        if (p.residual != nullptr) x.x += res[it].x, x.y += res[it].y, x.z += res[it].z, x.w += res[it].w;
        if ((er.mask_bits >> it) & 1u) x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.out_f32 != nullptr) *reinterpret_cast<float4*>(p.out_f32 + o) = x;
        if (p.out_hi != nullptr) {
            if (p.out_fmt != 0) {
                store_q8x4(x, p.out_hi, p.out_h8, p.out_l8, o);
            } else {
                uint32_t h0, l0, h1, l1;
                split_pack2(x.x, x.y, h0, l0);
                split_pack2(x.z, x.w, h1, l1);
                *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2(l0, l1);
            }
        }
    }
    __syncwarp();  // the staging buffer is rewritten by the next chunk
}

// One warp normalises complete output rows (fused LayerNorm of the CTA-pair kernel), TWO rows at a time so that the L2
// round trip of one row's loads overlaps the other row's reductions (a single row per iteration was latency-bound:
// ~1.7 k cycles per row, which stalled the epilogue warps long enough to starve the accumulator double buffer).
// Same arithmetic, in the same order, as layernorm_kernel (norm.cu): two-pass statistics held in registers, biased
// variance, eps 1e-5 — the values are bit-identical to the separate kernel's.
template <int V4>
__device__ __forceinline__ void ln_rows2(const GemmParams& p, size_t row_a, size_t row_b, bool has_b, int lane) {
    constexpr int D = V4 * 128;
    const float4* xa = reinterpret_cast<const float4*>(p.out_f32 + row_a * D);
    const float4* xb = reinterpret_cast<const float4*>(p.out_f32 + row_b * D);
    float4 va[V4], vb[V4];
#pragma unroll
    for (int i = 0; i < V4; ++i) va[i] = __ldcg(xa + lane + 32 * i);  // written by other SMs a moment ago: through L2
#pragma unroll
    for (int i = 0; i < V4; ++i) vb[i] = __ldcg(xb + lane + 32 * i);
    const float4* g4 = reinterpret_cast<const float4*>(p.ln_gamma);
    const float4* b4 = reinterpret_cast<const float4*>(p.ln_beta);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        if (which == 1 && !has_b) break;
        float4(&v)[V4] = which == 0 ? va : vb;
        const size_t row = which == 0 ? row_a : row_b;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
        for (int i = 0; i < V4; ++i) {
            const int c4 = lane + 32 * i;
            const float4 g = __ldg(g4 + c4), b = __ldg(b4 + c4);
            float4 y;
            y.x = (v[i].x - mean) * rstd * g.x + b.x;
            y.y = (v[i].y - mean) * rstd * g.y + b.y;
            y.z = (v[i].z - mean) * rstd * g.z + b.z;
            y.w = (v[i].w - mean) * rstd * g.w + b.w;
            if (p.ln_gelu) y.x = gelu_erf(y.x), y.y = gelu_erf(y.y), y.z = gelu_erf(y.z), y.w = gelu_erf(y.w);
            if (p.ln_out_f32 != nullptr) reinterpret_cast<float4*>(p.ln_out_f32 + row * D)[c4] = y;
            if (p.ln_out_hi != nullptr && p.out_fmt != 0) {
                store_q8x4(y, p.ln_out_hi, p.ln_out_h8, p.ln_out_l8, row * D + 4 * (size_t)c4);
            } else if (p.ln_out_hi != nullptr) {
                uint32_t h0, l0, h1, l1;
                split_pack2(y.x, y.y, h0, l0);
                split_pack2(y.z, y.w, h1, l1);
                reinterpret_cast<uint2*>(p.ln_out_hi + row * D)[c4] = make_uint2(h0, h1);
                reinterpret_cast<uint2*>(p.ln_out_lo + row * D)[c4] = make_uint2(l0, l1);
            }
        }
    }
}

// 16 of the 128 rows of a finished row block per epilogue warp. Not inlined: its registers (two rows) must not push
// the hot epilogue loop of the GEMM kernel into spills.
__device__ __noinline__ void ln_tile_rows(const GemmParams* pp, int batch, int row0, int ewarp, int lane) {
    const GemmParams& p = *pp;
    const int r_begin = ewarp * 16;
    for (int r = r_begin; r < r_begin + 16; r += 2) {
        const int ra = row0 + r, rb = ra + 1;
        if (ra >= p.rows_per_batch) break;
        const bool has_b = rb < p.rows_per_batch;
        const size_t ga = (size_t)batch * p.out_rows_per_batch + ra;
        const size_t gb = has_b ? ga + 1 : ga;
        switch (p.ldo) {
            case 512: ln_rows2<4>(p, ga, gb, has_b, lane); break;
            case 768: ln_rows2<6>(p, ga, gb, has_b, lane); break;
            case 1024: ln_rows2<8>(p, ga, gb, has_b, lane); break;
            default: ln_rows2<10>(p, ga, gb, has_b, lane); break;
        }
    }
}

template <int BLOCK_N, int BLOCK_K>
__global__ void __launch_bounds__(kThreads, 1) gemm_bf16x3_kernel(const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N, BLOCK_K>;
    constexpr int kBlockK = Cfg::kBlockK;
    constexpr int kATileBytes = Cfg::kATileBytes;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atoms (8 rows x 128 B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
    uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a_hi);
        tma_prefetch_desc(&p.a_lo);
        tma_prefetch_desc(&p.b_hi);
        tma_prefetch_desc(&p.b_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_base_slot, 2 * kAccCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below
    pdl_launch_dependents();

    const int tiles_m = p.batches * p.tiles_m_per_batch;
    const int num_tiles = tiles_m * p.n_tiles;
    const uint32_t b_tile_bytes = (uint32_t)p.umma_n * kBlockK * 2;
    const uint32_t stage_tx_bytes = 2u * kATileBytes + 2u * b_tile_bytes;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                const int mt = tile / p.n_tiles;
                const int batch = mt / p.tiles_m_per_batch;
                const int row0 = (mt - batch * p.tiles_m_per_batch) * kBlockM;
                const int a_k0 = p.a_k_per_ntile * n_tile;
                const int b_n0 = p.b_n_tiled ? n_tile * p.umma_n : 0;
                const int b_z0 = n_tile * p.b_z_per_ntile;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    const int kq = kb / p.kb_per_row;
                    const int kr = kb - kq * p.kb_per_row;
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    uint8_t* st = smem + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_tx_bytes);
                    const int a_row = row0 + kq * p.a_row_step + p.a_row_off;
                    tma_load_3d(st, &p.a_hi, &full_bar[stage], a_k0 + kr * kBlockK, a_row, batch);
                    tma_load_3d(st + kATileBytes, &p.a_lo, &full_bar[stage], a_k0 + kr * kBlockK, a_row, batch);
                    const int b_k = (p.b_k_linear ? kb : kr) * kBlockK;
                    const int b_z = p.b_k_linear ? 0 : b_z0 + kq;
                    tma_load_3d(st + 2 * kATileBytes, &p.b_hi, &full_bar[stage], b_k, b_n0, b_z);
                    tma_load_3d(st + 2 * kATileBytes + Cfg::kBTileBytes, &p.b_lo, &full_bar[stage], b_k, b_n0, b_z);
                    if (++stage == Cfg::kStages) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kBlockM, (uint32_t)p.umma_n);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t da_hi = make_smem_desc<kBlockK * 2>(st);
                    const uint64_t da_lo = make_smem_desc<kBlockK * 2>(st + kATileBytes);
                    const uint64_t db_hi = make_smem_desc<kBlockK * 2>(st + 2 * kATileBytes);
                    const uint64_t db_lo = make_smem_desc<kBlockK * 2>(st + 2 * kATileBytes + Cfg::kBTileBytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // advance 16 bf16 (= 32 bytes) along K inside the swizzle atom: +2 in (addr >> 4) units
                        const uint64_t ko = (uint64_t)(k * 2);
                        umma_bf16(d_tmem, da_lo + ko, db_hi + ko, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_bf16(d_tmem, da_hi + ko, db_lo + ko, idesc, 1u);
                        umma_bf16(d_tmem, da_hi + ko, db_hi + ko, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (++stage == Cfg::kStages) stage = 0, phase ^= 1u;
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
                if (++acc == 2) acc = 0, acc_phase ^= 1u;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;  // == warp % 4 : TMEM lane quadrant this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.n_tiles;
            const int mt = tile / p.n_tiles;
            const int batch = mt / p.tiles_m_per_batch;
            const int row0 = (mt - batch * p.tiles_m_per_batch) * kBlockM;
            const int row = row0 + ew * 32 + lane;
            const bool row_ok = row < p.rows_per_batch;
            const size_t m = (size_t)batch * p.out_rows_per_batch + (row_ok ? row : 0);
            const bool masked = (p.row_mask != nullptr) && row_ok && (p.row_mask[m] != 0);
            const int col0 = n_tile * p.umma_n;

            mbar_wait(&tmem_full[acc], acc_phase);
            __syncwarp();
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kAccCols;
            int c = 0;
            for (; c + 32 <= p.umma_n; c += 32) {
                uint32_t v[32];
                float4 res[8];
                load_residual<32>(p, row_ok, m, col0 + c, res);
                tmem_ld_32x32(t_row + (uint32_t)c, v);
                tmem_ld_wait();
                epilogue_cols<32>(p, v, row_ok, m, col0 + c, masked, res);
            }
            if (c < p.umma_n) {  // 16-column tail (umma_n % 32 == 16)
                uint32_t v[16];
                float4 res[4];
                load_residual<16>(p, row_ok, m, col0 + c, res);
                tmem_ld_32x16(t_row + (uint32_t)c, v);
                tmem_ld_wait();
                epilogue_cols<16>(p, v, row_ok, m, col0 + c, masked, res);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) acc = 0, acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * kAccCols);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) for 256-column tiles: one 256 x 256 output tile per cluster of two CTAs.
// Each CTA loads its own 128 rows of A (hi, lo) and 128 of the 256 W rows (hi, lo): 64 KB per 64-wide k-block
// instead of 96 KB, and every tcgen05.mma reads 8 KB instead of 12 KB of local shared memory — the single-CTA
// kernel is bound by shared-memory bandwidth (TMA writes + operand reads ~158 B/clk vs 128 B/clk available).
// Protocol: TMA of both CTAs credits the LEADER's full barrier; the leader issues M=256 MMAs and multicasts its
// commits to both CTAs' empty / tmem_full barriers; both epilogues (128 rows each) arrive on the leader's
// tmem_empty barrier.
// ------------------------------------------------------------------------------------------------
static constexpr int k2BlockK = 64;
static constexpr int k2TileBytes = 128 * k2BlockK * 2;    // 16 KB: A plane (128 rows) or half-B plane (128 rows)
static constexpr int k2StageBytes = 4 * k2TileBytes;      // A_hi A_lo B_hi B_lo
static constexpr int k2Stages = 3;
static constexpr int k2EpiWarps = 8;
static constexpr int k2EpiStageBytes = 32 * 32 * 4;  // per-warp transpose buffer of the coalesced epilogue
static constexpr int k2SmemBytes = k2Stages * k2StageBytes + 256 + k2EpiWarps * k2EpiStageBytes + 1024;

// 12 warps: TMA / MMA / TMEM-alloc / idle + EIGHT epilogue warps (two per TMEM lane quadrant, 128 columns each):
// with four, the GELU + split epilogue of a K=768 tile (fc1, QKV) took longer than its 12 k-blocks of MMAs.
static constexpr int k2Threads = 384;

// kScheme 0: bf16 hi/lo operands, 3 MMAs per product (k-blocks of 64). kScheme 1 ("f16q8"): per 128-wide k-block one
// stage of e4m3 correction operands (A_l8 | A_h8 | W_h8 | W_l8: two kind::f8f6f4 MMAs per 32-wide k-step) in a first
// pass over K, then one stage of fp16 operands (A16[k] | A16[k+64] | W16[k] | W16[k+64]: one kind::f16 MMA per 16-wide
// k-step) in a second pass — 16 instruction slots per 128 of K instead of 24. The corrections carry the factor
// 2^kQ8Scale; the first main MMA of the tile scales the accumulator back (scale-input-d).
template <int kScheme>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
    gemm2_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes);
    uint64_t* full_bar = bars;                   // [k2Stages] used in the leader only (count 2 + tx bytes)
    uint64_t* empty_bar = bars + k2Stages;       // [k2Stages] one per CTA, multicast commit
    uint64_t* tmem_full = bars + 2 * k2Stages;   // [2]        one per CTA, multicast commit
    uint64_t* tmem_empty = tmem_full + 2;        // [2]        leader only (16 epilogue warps of the pair)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const bool tracing = p.trace != nullptr && blockIdx.x == 0;
#define S3B_GTR(slot)                                        \
    do {                                                     \
        if (tracing) p.trace[slot] = (unsigned long long)clock64(); \
    } while (0)
    const long long t_entry = clock64();

    cluster_sync_all();  // both CTAs resident before the paired TMEM allocation
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&p.a_hi);
        tma_prefetch_desc(&p.a_lo);
        tma_prefetch_desc(&p.b_hi);
        tma_prefetch_desc(&p.b_lo);
    }
    if (warp == 1 && elect_one()) {
        for (int s = 0; s < k2Stages; ++s) {
            mbar_init(&full_bar[s], 2);  // leader's expect_tx arrive + the peer's remote arrive
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 16);  // 8 epilogue warps x 2 CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc_2cta(tmem_base_slot, 2 * kAccCols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    const long long t_prologue = clock64();
    pdl_wait();  // prologue overlapped the previous kernel's tail; global memory is touched only below
    pdl_launch_dependents();
    if (tracing && threadIdx.x == 0) {
        p.trace[0] = (unsigned long long)t_entry;
        p.trace[1] = (unsigned long long)t_prologue;
        p.trace[2] = (unsigned long long)clock64();
    }

    const int pairs_per_batch = (p.tiles_m_per_batch + 1) >> 1;
    const int num_pt = p.batches * pairs_per_batch * p.n_tiles;
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = cluster_id; pt < num_pt; pt += num_clusters) {
                const int n_tile = pt % p.n_tiles;
                const int mp = pt / p.n_tiles;
                const int batch = mp / pairs_per_batch;
                const int row0 = ((mp - batch * pairs_per_batch) * 2 + (int)rank) * kBlockM;
                const int a_k0 = p.a_k_per_ntile * n_tile;
                const int b_n0 = (p.b_n_tiled ? n_tile * p.umma_n : 0) + (int)rank * (p.umma_n >> 1);
                const int b_z0 = n_tile * p.b_z_per_ntile;
                if constexpr (kScheme == 0) {
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        const int kq = kb / p.kb_per_row;
                        const int kr = kb - kq * p.kb_per_row;
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* st = smem + stage * k2StageBytes;
                        // bytes of both CTAs: 2 x (A_hi + A_lo + two half-B planes of umma_n/2 rows)
                        if (leader)
                            mbar_arrive_expect_tx(&full_bar[stage],
                                                  2u * (2u * k2TileBytes + 2u * (uint32_t)(p.umma_n >> 1) * k2BlockK * 2u));
                        else mbar_arrive_remote(&full_bar[stage], 0);
                        const int a_row = row0 + kq * p.a_row_step + p.a_row_off;
                        tma_load_3d_2cta(st, &p.a_hi, &full_bar[stage], a_k0 + kr * k2BlockK, a_row, batch);
                        tma_load_3d_2cta(st + k2TileBytes, &p.a_lo, &full_bar[stage], a_k0 + kr * k2BlockK, a_row, batch);
                        const int b_k = (p.b_k_linear ? kb : kr) * k2BlockK;
                        const int b_z = p.b_k_linear ? 0 : b_z0 + kq;
                        tma_load_3d_2cta(st + 2 * k2TileBytes, &p.b_hi, &full_bar[stage], b_k, b_n0, b_z);
                        tma_load_3d_2cta(st + 3 * k2TileBytes, &p.b_lo, &full_bar[stage], b_k, b_n0, b_z);
                        if (++stage == k2Stages) stage = 0, phase ^= 1u;
                    }
                } else {
                    // pass 0: e4m3 correction operands, pass 1: fp16 main operands; 128 elements of K per stage
                    for (int pass = 0; pass < 2; ++pass) {
                        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                            const int kq = kb / p.kb_per_row;
                            const int kr = kb - kq * p.kb_per_row;
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* st = smem + stage * k2StageBytes;
                            // both passes move the same bytes: 128 rows x 128 B per A tile, umma_n/2 rows x 128 B per W tile
                            if (leader)
                                mbar_arrive_expect_tx(&full_bar[stage],
                                                      2u * (2u * k2TileBytes + 2u * (uint32_t)(p.umma_n >> 1) * 128u));
                            else mbar_arrive_remote(&full_bar[stage], 0);
                            const int a_row = row0 + kq * p.a_row_step + p.a_row_off;
                            const int a_k = a_k0 + kr * 128;
                            const int b_k = (p.b_k_linear ? kb : kr) * 128;
                            const int b_z = p.b_k_linear ? 0 : b_z0 + kq;
                            if (pass == 0) {
                                tma_load_3d_2cta(st, &p.a_l8, &full_bar[stage], a_k, a_row, batch);
                                tma_load_3d_2cta(st + k2TileBytes, &p.a_h8, &full_bar[stage], a_k, a_row, batch);
                                tma_load_3d_2cta(st + 2 * k2TileBytes, &p.b_h8, &full_bar[stage], b_k, b_n0, b_z);
                                tma_load_3d_2cta(st + 3 * k2TileBytes, &p.b_l8, &full_bar[stage], b_k, b_n0, b_z);
                            } else {
                                tma_load_3d_2cta(st, &p.a_hi, &full_bar[stage], a_k, a_row, batch);
                                tma_load_3d_2cta(st + k2TileBytes, &p.a_hi, &full_bar[stage], a_k + 64, a_row, batch);
                                tma_load_3d_2cta(st + 2 * k2TileBytes, &p.b_hi, &full_bar[stage], b_k, b_n0, b_z);
                                tma_load_3d_2cta(st + 3 * k2TileBytes, &p.b_hi, &full_bar[stage], b_k + 64, b_n0, b_z);
                            }
                            if (++stage == k2Stages) stage = 0, phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (elect_one()) {
            const uint32_t idesc = kScheme == 0 ? make_idesc_bf16(2 * kBlockM, (uint32_t)p.umma_n)
                                                : make_idesc_f16(2 * kBlockM, (uint32_t)p.umma_n);
            const int k_steps = p.k_steps > 0 ? p.k_steps : k2BlockK / 16;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int done = 0;
            for (int pt = cluster_id; pt < num_pt; pt += num_clusters, ++done) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * kAccCols;
                if constexpr (kScheme == 0) {
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        if (done == 0 && kb == 0) S3B_GTR(3);
                        const uint32_t st = smem_u32(smem + stage * k2StageBytes);
                        const uint64_t da_hi = make_smem_desc<128>(st);
                        const uint64_t da_lo = make_smem_desc<128>(st + k2TileBytes);
                        const uint64_t db_hi = make_smem_desc<128>(st + 2 * k2TileBytes);
                        const uint64_t db_lo = make_smem_desc<128>(st + 3 * k2TileBytes);
#pragma unroll
                        for (int k = 0; k < k2BlockK / 16; ++k) {
                            if (k < k_steps) {
                                const uint64_t ko = (uint64_t)(k * 2);
                                umma_bf16_2cta(d_tmem, da_lo + ko, db_hi + ko, idesc, (kb | k) != 0 ? 1u : 0u);
                                umma_bf16_2cta(d_tmem, da_hi + ko, db_lo + ko, idesc, 1u);
                                umma_bf16_2cta(d_tmem, da_hi + ko, db_hi + ko, idesc, 1u);
                            }
                        }
                        umma_commit_2cta(&empty_bar[stage]);
                        if (++stage == k2Stages) stage = 0, phase ^= 1u;
                    }
                } else {
                    for (int pass = 0; pass < 2; ++pass) {
                        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            if (done == 0 && kb == 0 && pass == 0) S3B_GTR(3);
                            const uint32_t st = smem_u32(smem + stage * k2StageBytes);
                            // pass 0: A_l8 x W_h8 and A_h8 x W_l8 (K = 32 per MMA, 32 bytes);
                            // pass 1: A16[k..k+64) x W16[k..k+64) and the next 64 (K = 16 per MMA, 32 bytes)
                            const uint64_t da0 = make_smem_desc<128>(st), da1 = make_smem_desc<128>(st + k2TileBytes);
                            const uint64_t db0 = make_smem_desc<128>(st + 2 * k2TileBytes);
                            const uint64_t db1 = make_smem_desc<128>(st + 3 * k2TileBytes);
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t ko = (uint64_t)(k * 2);
                                    const uint64_t da = (t == 0 ? da0 : da1) + ko, db = (t == 0 ? db0 : db1) + ko;
                                    if (pass == 0) {
                                        if (p.q8_debug == 1) umma_bf16_2cta(d_tmem, da, db, idesc, (kb | t | k) != 0 ? 1u : 0u);
                                        else if (p.q8_debug != 3) umma_q8_2cta(d_tmem, da, db, idesc, (kb | t | k) != 0 ? 1u : 0u);
                                    } else if ((kb | t | k) == 0 && p.q8_debug != 2 && p.q8_debug != 3) {
                                        umma_f16_2cta_scaled(d_tmem, da, db, idesc);  // D = A*B + D * 2^-kQ8Scale
                                    } else if (p.q8_debug != 4 || (kb | t | k) == 0) {
                                        umma_bf16_2cta(d_tmem, da, db, idesc, (p.q8_debug == 3 && (kb | t | k) == 0) ? 0u : 1u);
                                    }
                                }
                            }
                            umma_commit_2cta(&empty_bar[stage]);
                            if (++stage == k2Stages) stage = 0, phase ^= 1u;
                        }
                    }
                }
                umma_commit_2cta(&tmem_full[acc]);
                S3B_GTR(4);
                if (++acc == 2) acc = 0, acc_phase ^= 1u;
            }
            // the peer's epilogue arrives remotely on this CTA's tmem_empty barriers: drain before teardown
            if (done > 0) {
                const int last = done - 1;
                mbar_wait(&tmem_empty[last & 1], (uint32_t)((last >> 1) & 1));
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, 128 rows each) =====================
        const int ew = (warp - 4) & 3;        // TMEM lane quadrant (== warp % 4)
        const int chalf = (warp - 4) >> 2;    // which 128-column half of the accumulator this warp drains
        uint8_t* epi_stage = smem + k2Stages * k2StageBytes + 256 + (warp - 4) * k2EpiStageBytes;
        int acc = 0;
        uint32_t acc_phase = 0;
        const bool etr = tracing && warp == 4 && lane == 0;
        int ntile_done = 0;
        for (int pt = cluster_id; pt < num_pt; pt += num_clusters, ++ntile_done) {
            const int n_tile = pt % p.n_tiles;
            const int mp = pt / p.n_tiles;
            const int batch = mp / pairs_per_batch;
            const int row0 = ((mp - batch * pairs_per_batch) * 2 + (int)rank) * kBlockM;
            const int row = row0 + ew * 32 + lane;
            const bool row_ok = row < p.rows_per_batch;
            const size_t m = (size_t)batch * p.out_rows_per_batch + (row_ok ? row : 0);
            const bool masked = (p.row_mask != nullptr) && row_ok && (p.row_mask[m] != 0);
            const int col0 = n_tile * p.umma_n;

            mbar_wait(&tmem_full[acc], acc_phase);
            __syncwarp();
            tc_fence_after();
            if (etr) p.trace[ntile_done == 0 ? 5 : 7] = (unsigned long long)clock64();
            const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc * kAccCols;
            const int chw = p.umma_n >> 1;  // columns per epilogue warp group (128 or 64)
            EpiRows er;
            epi_rows_init(p, er, batch, row0 + ew * 32, lane, masked);
            float4 res_next[8];
            epi_load_residual(p, er, col0 + chalf * chw, lane, res_next);
            for (int c = chalf * chw; c < (chalf + 1) * chw; c += 32) {
                uint32_t v[32];
                float4 res[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) res[j] = res_next[j];
                tmem_ld_32x32(t_row + (uint32_t)c, v);
                if (c + 32 < (chalf + 1) * chw) epi_load_residual(p, er, col0 + c + 32, lane, res_next);
                tmem_ld_wait();
                if (p.qkv_mode && col0 + c >= 2 * p.D) {
                    float4 none[8];
                    epilogue_cols<32>(p, v, row_ok, m, col0 + c, masked, none);
                } else {
                    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.bias != nullptr) bias = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + c) + (lane & 7));
                    epi_chunk_coalesced(p, er, v, epi_stage, col0 + c, lane, res, bias);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tmem_empty[acc]);
                else mbar_arrive_remote(&tmem_empty[acc], 0);
            }
#ifdef S3B_ENABLE_FUSED_LN
            // Compiled out by default: the call below costs the epilogue loop ~50 registers of caller-saved state
            // (ptxas then keeps the residual prefetch array in local memory), which slowed EVERY GEMM by ~6 %, more than
            // the fusion can win back (measured, profiles/README.md r2d / r2f). The code is kept for the experiment.
            if (p.ln_gamma != nullptr) {
                // ---- fused LayerNorm: the last CTA to finish its n-tile of these 128 rows normalises them ------------
                // (all 8 epilogue warps of this CTA take part; named barrier 1 = the epilogue warps only)
                uint32_t* last_flag = reinterpret_cast<uint32_t*>(smem + k2Stages * k2StageBytes + 192);
                asm volatile("bar.sync 1, 256;" ::: "memory");  // every warp's stores of this tile are issued
                if (warp == 4 && lane == 0) {
                    __threadfence();  // release: this CTA's part of the rows is visible device-wide
                    const int rb = batch * p.tiles_m_per_batch + row0 / kBlockM;
                    const unsigned int old = atomicAdd(p.ln_counter + rb, 1u);
                    const bool last = old == (unsigned int)(p.n_tiles - 1);
                    if (last) p.ln_counter[rb] = 0;  // ready for the next launch that uses this counter array
                    *last_flag = last ? 1u : 0u;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (*last_flag != 0u) {
                    __threadfence();  // acquire: the other CTAs' stores
                    ln_tile_rows(&p, batch, row0, warp - 4, lane);
                }
            }
#endif
            if (etr) p.trace[ntile_done == 0 ? 6 : 8] = (unsigned long long)clock64();
            if (++acc == 2) acc = 0, acc_phase ^= 1u;
        }
        if (etr) p.trace[10] = (unsigned long long)ntile_done;
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 2 * kAccCols);
    }
    if (threadIdx.x == 0) S3B_GTR(9);
#undef S3B_GTR
}

static cudaError_t launch_pair(const GemmParams& p, int sm_count, cudaStream_t stream) {
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.current();
    if (!attr_set) {
        cudaError_t e =
            cudaFuncSetAttribute(gemm2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    // umma_n / 2 columns per epilogue warp group, in 32-column chunks
    if (p.umma_n != 256 && p.umma_n != 192 && p.umma_n != 128) return cudaErrorInvalidValue;
    if (p.scheme == 0 ? p.block_k != k2BlockK : (p.block_k != 128 || p.umma_n == 192 || p.k_steps != 0))
        return cudaErrorInvalidValue;
    if (p.out_fmt != 0 && (p.qkv_mode || (p.out_hi != nullptr && (p.out_h8 == nullptr || p.out_l8 == nullptr))))
        return cudaErrorInvalidValue;
#ifndef S3B_ENABLE_FUSED_LN
    if (p.ln_gamma != nullptr) return cudaErrorNotSupported;
#endif
    if (p.ln_gamma != nullptr &&
        (p.out_f32 == nullptr || p.ln_counter == nullptr || p.qkv_mode || p.n_tiles * p.umma_n != p.ldo ||
         (p.ldo != 512 && p.ldo != 768 && p.ldo != 1024 && p.ldo != 1280)))
        return cudaErrorInvalidValue;
    const int num_pt = p.batches * ((p.tiles_m_per_batch + 1) / 2) * p.n_tiles;
    if (num_pt <= 0) return cudaSuccess;
    const int clusters = num_pt < sm_count / 2 ? num_pt : sm_count / 2;
    if (p.scheme != 0) return launch_pdl(gemm2_kernel<1>, dim3(2 * clusters), dim3(k2Threads), k2SmemBytes, stream, p);
    return launch_pdl(gemm2_kernel<0>, dim3(2 * clusters), dim3(k2Threads), k2SmemBytes, stream, p);
}

template <int BLOCK_N, int BLOCK_K>
static cudaError_t launch_impl(const GemmParams& p, int sm_count, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N, BLOCK_K>;
    static PerDeviceOnce attr_once;
    bool& attr_set = attr_once.current();
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16x3_kernel<BLOCK_N, BLOCK_K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int num_tiles = p.batches * p.tiles_m_per_batch * p.n_tiles;
    if (num_tiles <= 0) return cudaSuccess;
    const int grid = num_tiles < sm_count ? num_tiles : sm_count;
    return launch_pdl(gemm_bf16x3_kernel<BLOCK_N, BLOCK_K>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, p);
}

cudaError_t launch_gemm_bf16x3(const GemmParams& p, int sm_count, cudaStream_t stream) {
    if (p.umma_n % 16 != 0 || p.umma_n < 16 || p.umma_n > 256) return cudaErrorInvalidValue;
    if (p.two_cta) return launch_pair(p, sm_count, stream);
    if (p.block_k != gemm_block_k(p.umma_n)) return cudaErrorInvalidValue;
    if (p.umma_n > 128) return launch_impl<256, 32>(p, sm_count, stream);
    if (p.umma_n > 64) return launch_impl<128, 64>(p, sm_count, stream);
    return launch_impl<64, 64>(p, sm_count, stream);
}

// ------------------------------------------------------------------------------------------------
// tensor-map encoding through the driver entry point (no link-time libcuda dependency: the C-ABI
// library must load on a GPU-less host for the symbol-export test)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

int encode_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                        uint64_t stride2, uint32_t box0, uint32_t box1) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) return -1;
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1 * 2, stride2 * 2};  // bytes
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    // the swizzle span equals the box row: 64 elements -> 128 B, 32 elements -> 64 B
    if (box0 != 64 && box0 != 32) return -2;
    const CUtensorMapSwizzle sw = box0 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}


int encode_tmap_u8_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                      uint64_t stride2, uint32_t box0, uint32_t box1) {
    PFN_encodeTiled fn = get_encode_fn();
    if (fn == nullptr) return -1;
    if (box0 != 128) return -2;  // one 128-byte swizzle row of e4m3 elements
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1, stride2};  // bytes == elements
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

}  // namespace s3b
