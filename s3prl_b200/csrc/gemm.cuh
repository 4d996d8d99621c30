// Host-visible description of one error-compensated (bf16 hi/lo, 3-MMA) tcgen05 GEMM launch.
// One kernel serves every contraction on the hot path: the strided conv stack
// (conv 1..6 as implicit GEMM on channels-last activations), post_extract_proj, the grouped positional
// conv (one k-block per tap, time shift expressed in the TMA row coordinate), QKV, out_proj, fc1, fc2.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace s3b {

struct GemmParams {
    // A operand (activations): bf16 hi / lo, 3-D tensor maps (k, row, batch), box {64, 128, 1}, SWIZZLE_128B
    CUtensorMap a_hi, a_lo;
    // B operand (weights, [N][K] K-major): bf16 hi / lo, 3-D tensor maps (k, n, z), box {64, umma_n, 1}
    CUtensorMap b_hi, b_lo;

    // scheme 1 ("f16q8", CTA-pair kernel only): fp16 main product + two e4m3 correction products.
    //   a_hi / b_hi are the fp16 planes (box {64, rows}), a_h8 / a_l8 / b_h8 / b_l8 the e4m3 planes (box {128, rows});
    //   k-blocks are 128 elements wide: block_k == 128, K % 128 == 0; a_lo / b_lo are unused.
    int scheme;
    int q8_debug;  // s3b_gemm_bench only (timing experiments, wrong numerics): 1 = corrections issued as kind::f16,
                   // 2 = no scale-input-d, 3 = correction MMAs skipped, 4 = main MMAs skipped
    CUtensorMap a_h8, a_l8, b_h8, b_l8;

    // ---- tiling -------------------------------------------------------------------------------
    int batches;            // extent of A's 3rd dim that is tiled over
    int rows_per_batch;     // valid output rows per batch
    int tiles_m_per_batch;  // ceil(rows_per_batch / 128)
    int n_tiles;            // output column tiles
    int umma_n;             // columns per tile (multiple of 16, <= 256)
    int two_cta;            // 1: CTA-pair kernel (umma_n == 256, block_k == 64, B boxes hold umma_n/2 rows)
    int block_k;            // bf16 elements per k-block: gemm_block_k(umma_n) (32 for 128x256 tiles, else 64)
    int num_k_blocks;       // k-blocks per output tile
    // k-block kb reads   A box at (a_k_per_ntile*n_tile + (kb % kb_per_row)*block_k,
    //                              row0 + (kb / kb_per_row)*a_row_step + a_row_off, batch)
    //                    B box at (b_k_linear ? kb*block_k : (kb % kb_per_row)*block_k, b_n_tiled ? n_tile*umma_n : 0,
    //                              b_k_linear ? 0 : kb / kb_per_row + n_tile*b_z_per_ntile)
    int kb_per_row;
    int a_row_step;
    int a_row_off;
    int a_k_per_ntile;
    int b_n_tiled;
    int b_k_linear;
    int b_z_per_ntile;

    // ---- epilogue: v = acc (+bias[n]) ; gelu? ; (+residual[m][n]) ; row_mask[m] ? 0 -----------------
    int out_rows_per_batch;  // flat output row m = batch*out_rows_per_batch + row
    int ldo;                 // leading dimension (elements) of out_f32/out_hi/out_lo/residual
    const float* bias;
    const float* residual;
    const uint8_t* row_mask;
    int gelu;
    float* out_f32;
    __nv_bfloat16* out_hi;
    __nv_bfloat16* out_lo;
    // out_fmt 1: out_hi is an fp16 plane and out_h8 / out_l8 the two e4m3 planes of the f16q8 operand format (same for
    // the fused LayerNorm outputs); out_fmt 0: out_hi / out_lo are the bf16 hi / lo planes.
    int out_fmt;
    uint8_t* out_h8;
    uint8_t* out_l8;
    float* out_pre;  // optional: v BEFORE the residual add (fc2 output, "fairseq_layers_before_residual"), same layout

    // ---- epilogue, QKV scatter mode (qkv_mode != 0): columns [0,D) -> q*scale, [D,2D) -> k, [2D,3D) -> v
    // q,k: [B][H][T][64] split bf16 ; v transposed: [B][H][64][Tp] split bf16. Flat row m = b*T + t.
    int qkv_mode;
    int T, Tp, H, D;
    float q_scale;
    __nv_bfloat16 *q_hi, *q_lo, *k_hi, *k_lo, *vt_hi, *vt_lo;

    // ---- optional fused LayerNorm over the whole output row (CTA-pair kernel, N == ldo == 128 * ln_v4) ------------
    // The n-tiles of a 128-row block are finished by different CTAs; each counts itself in on ln_counter[row block]
    // after its stores, and the LAST one to arrive normalises the 128 complete rows of out_f32 (re-read from L2):
    //   y = LayerNorm(out_f32 row) * ln_gamma + ln_beta (eps 1e-5), optional GELU -> ln_out_f32 / ln_out_hi,lo
    // The values are the ones the separate layernorm_kernel would produce, bit for bit (same per-row arithmetic), and
    // the LayerNorm traffic overlaps the MMAs of the CTA's next tile instead of running as its own HBM-bound kernel.
    const float* ln_gamma;  // null = no fused LayerNorm
    const float* ln_beta;
    int ln_gelu;
    float* ln_out_f32;
    __nv_bfloat16* ln_out_hi;
    __nv_bfloat16* ln_out_lo;
    uint8_t* ln_out_h8;
    uint8_t* ln_out_l8;
    unsigned int* ln_counter;  // [batches * tiles_m_per_batch], zero on entry, left zero

    double alg_flops;  // host-side accounting only: 2*M*N*K with the un-padded K

    // CTA-pair kernel: 16-wide MMA k-steps issued per 64-wide k-block (0 = all 4). pos_conv with 48 channels per group
    // loads 64-wide boxes but only the first 3 k-steps carry non-zero weights.
    int k_steps;

    // debug timeline (tools/gemm_trace.py): 16 clock64 stamps of CTA 0, or nullptr
    unsigned long long* trace;
};

// k-block width the kernel instantiation for this tile width uses (tensor-map boxes must match)
inline int gemm_block_k(int umma_n) { return umma_n > 128 ? 32 : 64; }

// Launch on `stream`. Returns cudaGetLastError() of the launch.
cudaError_t launch_gemm_bf16x3(const GemmParams& p, int sm_count, cudaStream_t stream);

// Encode a 3-D bf16 tensor map with the 128B swizzle. strides in ELEMENTS for dims 1 and 2.
// Returns 0 on success, else a CUresult / -1 (driver entry point missing).
int encode_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                        uint64_t stride2, uint32_t box0, uint32_t box1);
// Same for a 1-byte (e4m3) plane: box0 must be 128 elements (one 128-byte swizzle row).
int encode_tmap_u8_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                      uint64_t stride2, uint32_t box0, uint32_t box1);

}  // namespace s3b
