// Launchers of the non-GEMM kernels (frontend.cu, norm.cu, attention.cu, fbank.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace s3b {

// ---- frontend.cu ---------------------------------------------------------------------------------
cudaError_t launch_split(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t s);
cudaError_t launch_wav_pack(const float* const* wavs, const long long* lens, int B, long long Lpad, int normalize,
                            float* mean_rstd_ws, float* out, cudaStream_t s);
size_t conv0_ws_part_floats(int B, int L0);
cudaError_t launch_split_q8(const float* x, __nv_bfloat16* p16, uint8_t* h8, uint8_t* l8, size_t n, int weight,
                            cudaStream_t s);
cudaError_t launch_conv0_groupnorm(const float* x, int B, long long L, int L0, const float* w, const float* gamma,
                                   const float* beta, float* ws_part, float* ws_scale_shift, const OutPlanes& op,
                                   cudaStream_t s);
cudaError_t launch_conv0_layernorm(const float* x, int B, long long L, int L0, const float* w, const float* cbias,
                                   const float* gamma, const float* beta, const OutPlanes& op, cudaStream_t s);

// ---- norm.cu -------------------------------------------------------------------------------------
// y = LayerNorm_D(x) * gamma + beta (eps 1e-5, biased var), optional GELU; D in {512, 768, 1024}.
// Any of out_f32 / (out_hi,out_lo) may be null. x and out_f32 may alias.
cudaError_t launch_layernorm(const float* x, size_t M, int D, const float* gamma, const float* beta, int gelu,
                             float* out_f32, const OutPlanes& op, cudaStream_t s);
// out[m] = sum_l w[l] * hs[l][m]   (Featurizer._weighted_sum, interfaces.py:217-248; w already softmaxed)
cudaError_t launch_posconv_combine(const float* z, const float* x, const float* bias, int B, int T, int D, int cpg,
                                   const float* gamma, const float* beta, int do_ln, int mode, float* out_f32,
                                   const OutPlanes& op, cudaStream_t s);
cudaError_t launch_weighted_sum(const float* hs, int NL, size_t n_per_layer, const float* w, float* out,
                                cudaStream_t s);
// grad_w[l] = sum_m hs[l][m] * gout[m]
cudaError_t launch_weighted_sum_bwd(const float* hs, int NL, size_t n_per_layer, const float* gout, float* grad_w,
                                    cudaStream_t s);

// ---- peer.cu: fused weighted sum + push all-gather over peer memory --------------------------------------------------
cudaError_t launch_weighted_sum_push(const float* hs, int NL, size_t n_per_layer, size_t layer_stride, const float* w,
                                     float* const* peer_dst, uint32_t* const* peer_flag, int n_peers, uint32_t seq,
                                     unsigned int* counter, cudaStream_t s);
cudaError_t launch_wait_flags(const uint32_t* flags, int n, uint32_t seq, cudaStream_t s);

// ---- attention.cu --------------------------------------------------------------------------------
struct AttnParams {
    CUtensorMap q_hi, q_lo;    // [B*H][T][64]  box {64, 128, 1}
    CUtensorMap k_hi, k_lo;    // [B*H][T][64]  box {64,  64, 1}
    CUtensorMap vt_hi, vt_lo;  // [B*H][64][Tp] (dim0 = T valid keys) box {64, 64, 1}
    int B, H, T, D;
    const int* kv_len;          // [B] valid (un-padded) key count, >= 1
    // optional WavLM gated relative position bias: S[q][k] += gate[b][h][q] * table[h][k - q + bias_center]
    const float* bias_table;    // [H][bias_stride], entry (h, k - q + bias_center); or null
    int bias_stride, bias_center;
    const float* gate;          // [B][H][T] or null
    OutPlanes ctx;  // [B*T][D] A operand of out_proj, in either operand format
    // optional timeline (debug): clock64 stamps of CTA `trace_block`, [2 roles][16 blocks][8 slots]
    long long* trace;
    int trace_block;
};
cudaError_t launch_attention(const AttnParams& p, cudaStream_t s);

}  // namespace s3b
