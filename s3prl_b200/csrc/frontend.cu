// HBM-bound front-end kernels of the conv feature extractor and the small glue ops:
//   * waveform packing (+ optional per-utterance normalisation)       hubert/expert.py:56-66
//   * conv-0 (Conv1d(1,512,10,5)) fused with GroupNorm(512,512) + GELU (extractor_mode "default")
//     or per-frame LayerNorm(512) + GELU (extractor_mode "layer_norm") wav2vec2_model.py:2869-2934
//   * fp32 -> bf16 hi/lo split (weights at load time)
// Outputs are channels-last [B][L0][512] bf16 hi/lo so that conv-1 is a plain K-major GEMM operand.
#include "common.cuh"
#include "kernels.cuh"

namespace s3b {

// ------------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) bf16 split
// ------------------------------------------------------------------------------------------------
__global__ void split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        __nv_bfloat16 h, l;
        split_bf16(x[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

cudaError_t launch_split(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const int threads = 256;
    size_t blocks = (n + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_kernel<<<(unsigned)blocks, threads, 0, s>>>(x, hi, lo, n);
    return cudaGetLastError();
}

// fp32 -> f16q8 planes (common.cuh): activations (weight == 0) or weights (weight != 0: wh8 = e4m3(w * 2^ShiftB),
// wl8 = e4m3((w - w16) * 2^ShiftD))
__global__ void split_q8_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ p16,
                                uint8_t* __restrict__ h8, uint8_t* __restrict__ l8, size_t n2, int weight) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) {
        const float a = x[2 * i], b = x[2 * i + 1];
        uint32_t h;
        uint16_t q_h, q_l;
        if (weight == 0) {
            split_q8_pack2(a, b, h, q_h, q_l);
        } else {
            asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
            float ha, hb;
            asm("{ .reg .b16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi; }"
                : "=f"(ha), "=f"(hb)
                : "r"(h));
            const float sb = (float)(1 << kQ8ShiftB), sd = (float)(1 << kQ8ShiftD);
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(q_h) : "f"(b * sb), "f"(a * sb));
            asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(q_l) : "f"((b - hb) * sd), "f"((a - ha) * sd));
        }
        reinterpret_cast<uint32_t*>(p16)[i] = h;
        reinterpret_cast<uint16_t*>(h8)[i] = q_h;
        reinterpret_cast<uint16_t*>(l8)[i] = q_l;
    }
}

cudaError_t launch_split_q8(const float* x, __nv_bfloat16* p16, uint8_t* h8, uint8_t* l8, size_t n, int weight,
                            cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    if (n & 1) return cudaErrorInvalidValue;
    const size_t n2 = n / 2;
    size_t blocks = (n2 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    split_q8_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, p16, h8, l8, n2, weight);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// pack B ragged waveforms into a zero-padded [B][Lpad] buffer; optional F.layer_norm(wav, wav.shape)
// (eps 1e-5, biased variance) per utterance (task_cfg.normalize, hubert/expert.py:57-58)
// ------------------------------------------------------------------------------------------------
__global__ void wav_stats_kernel(const float* const* __restrict__ wavs, const long long* __restrict__ lens,
                                 float* __restrict__ mean_rstd) {
    // one block per utterance; two-pass (mean, then centred second moment), fp32 with double block combine
    const int b = blockIdx.x;
    const float* w = wavs[b];
    const long long n = lens[b];
    __shared__ double red[32];
    __shared__ float s_mean;
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += (double)w[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        s_mean = (float)(t / (double)n);
    }
    __syncthreads();
    const float mean = s_mean;
    acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = w[i] - mean;
        acc += (double)(d * d);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        mean_rstd[2 * b] = mean;
        mean_rstd[2 * b + 1] = rsqrtf((float)(t / (double)n) + 1e-5f);
    }
}

__global__ void wav_pack_kernel(const float* const* __restrict__ wavs, const long long* __restrict__ lens,
                                const float* __restrict__ mean_rstd, float* __restrict__ out, long long Lpad) {
    const int b = blockIdx.y;
    const float* w = wavs[b];
    const long long n = lens[b];
    float mean = 0.f, rstd = 1.f;
    if (mean_rstd != nullptr) mean = mean_rstd[2 * b], rstd = mean_rstd[2 * b + 1];
    float* o = out + (size_t)b * Lpad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Lpad;
         i += (long long)gridDim.x * blockDim.x)
        o[i] = (i < n) ? (w[i] - mean) * rstd : 0.0f;
}

cudaError_t launch_wav_pack(const float* const* wavs, const long long* lens, int B, long long Lpad, int normalize,
                            float* mean_rstd_ws, float* out, cudaStream_t s) {
    if (normalize) wav_stats_kernel<<<B, 1024, 0, s>>>(wavs, lens, mean_rstd_ws);
    dim3 grid((unsigned)((Lpad + 1023) / 1024 < 64 ? (Lpad + 1023) / 1024 : 64), B);
    wav_pack_kernel<<<grid, 256, 0, s>>>(wavs, lens, normalize ? mean_rstd_ws : nullptr, out, Lpad);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// conv-0: z[b][t][c] = sum_{j<10} w[c][j] * x[b][5t+j]      (no bias in "default" mode)
// 256 threads, thread = channel pair (2*tid, 2*tid+1), block = TCH consecutive frames of one utterance.
// ------------------------------------------------------------------------------------------------
static constexpr int kC0 = 512;
static constexpr int kC0K = 10;
static constexpr int kC0S = 5;
static constexpr int kTCH = 256;                        // frames per block
static constexpr int kXTile = kTCH * kC0S + 8;          // samples staged per block (+ halo, padded)

__device__ __forceinline__ void conv0_load_tile(const float* __restrict__ xb, long long L, long long s0, float* xs) {
    for (int i = threadIdx.x; i < kXTile; i += blockDim.x) {
        const long long g = s0 + i;
        xs[i] = (g < L) ? xb[g] : 0.0f;
    }
}

// 4 consecutive frames x 2 channels from 25 staged samples
__device__ __forceinline__ void conv0_quad(const float* xs, int tl, const float (&w0)[kC0K], const float (&w1)[kC0K],
                                           float (&z0)[4], float (&z1)[4]) {
    float xv[28];
    const float4* x4 = reinterpret_cast<const float4*>(xs + tl * kC0S);  // tl % 4 == 0 -> 16B aligned
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const float4 v = x4[i];
        xv[4 * i] = v.x, xv[4 * i + 1] = v.y, xv[4 * i + 2] = v.z, xv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < kC0K; ++j) {
            a0 = fmaf(w0[j], xv[q * kC0S + j], a0);
            a1 = fmaf(w1[j], xv[q * kC0S + j], a1);
        }
        z0[q] = a0, z1[q] = a1;
    }
}

// pass 1: second-order moments of the 10-sample conv windows. With z_t[c] = sum_j w[c][j] x[5t+j],
//   sum_t z_t[c]   = sum_j w[c][j] m_j,            m_j  = sum_t x[5t+j]
//   sum_t z_t[c]^2 = sum_jk w[c][j] w[c][k] R_jk,  R_jk = sum_t x[5t+j] x[5t+k]
// so the GroupNorm statistics of all 512 channels follow from 10 + 55 numbers per utterance: one cheap pass over
// the waveform instead of a full conv pass (the previous conv0_stats_kernel took 0.29 ms at C2).
static constexpr int kNMom = kC0K + kC0K * (kC0K + 1) / 2;  // 65

__global__ void __launch_bounds__(256) conv0_moments_kernel(const float* __restrict__ x, long long L, int L0,
                                                            float* __restrict__ part /*[B][nchunk][65]*/,
                                                            int nchunk) {
    __shared__ __align__(16) float xs[kXTile];
    __shared__ float red[8][kNMom];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int t0 = chunk * kTCH;
    conv0_load_tile(x + (size_t)b * L, L, (long long)t0 * kC0S, xs);
    __syncthreads();
    const int t = t0 + threadIdx.x;
    float xv[kC0K];
#pragma unroll
    for (int j = 0; j < kC0K; ++j) xv[j] = (t < L0) ? xs[threadIdx.x * kC0S + j] : 0.0f;
    float mom[kNMom];
    int idx = 0;
#pragma unroll
    for (int j = 0; j < kC0K; ++j) mom[idx++] = xv[j];
#pragma unroll
    for (int j = 0; j < kC0K; ++j)
#pragma unroll
        for (int k = j; k < kC0K; ++k) mom[idx++] = xv[j] * xv[k];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kNMom; ++i) {
        const float v = warp_sum(mom[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNMom) {
        float v = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
        part[((size_t)b * nchunk + chunk) * kNMom + threadIdx.x] = v;
    }
}

// finalize: GroupNorm(512 groups == per channel) statistics over ALL L0 frames of the padded batch
// (Fp32GroupNorm, eps 1e-5, biased variance; wav2vec2_model.py:1841-1853, 2898-2904), in double:
//   y = (z - mean) * rstd * gamma + beta  =  z * scale + shift
__global__ void __launch_bounds__(kC0) conv0_finalize_kernel(const float* __restrict__ part, int nchunk, int L0,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             float* __restrict__ scale, float* __restrict__ shift) {
    __shared__ double mom[kNMom];
    const int b = blockIdx.x, c = threadIdx.x;
    if (c < kNMom) {
        double acc = 0.0;
        for (int k = 0; k < nchunk; ++k) acc += (double)part[((size_t)b * nchunk + k) * kNMom + c];
        mom[c] = acc;
    }
    __syncthreads();
    double wc[kC0K];
#pragma unroll
    for (int j = 0; j < kC0K; ++j) wc[j] = (double)w[c * kC0K + j];
    double s = 0.0, q = 0.0;
    int idx = kC0K;
#pragma unroll
    for (int j = 0; j < kC0K; ++j) {
        s += wc[j] * mom[j];
#pragma unroll
        for (int k = j; k < kC0K; ++k) q += (k == j ? 1.0 : 2.0) * wc[j] * wc[k] * mom[idx++];
    }
    const double mean = s / (double)L0;
    double var = q / (double)L0 - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float g = gamma[c] * rstd;
    scale[b * kC0 + c] = g;
    shift[b * kC0 + c] = beta[c] - (float)mean * g;
}

// pass 2 ("default" mode): recompute z, GroupNorm affine, GELU, split, channels-last store
__global__ void __launch_bounds__(256) conv0_apply_gn_kernel(const float* __restrict__ x, long long L, int L0,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ scale,
                                                             const float* __restrict__ shift, const OutPlanes op) {
    __shared__ __align__(16) float xs[kXTile];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kTCH;
    conv0_load_tile(x + (size_t)b * L, L, (long long)t0 * kC0S, xs);
    float w0[kC0K], w1[kC0K];
    const int c0 = 2 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < kC0K; ++j) w0[j] = w[c0 * kC0K + j], w1[j] = w[(c0 + 1) * kC0K + j];
    const float sc0 = scale[b * kC0 + c0], sc1 = scale[b * kC0 + c0 + 1];
    const float sh0 = shift[b * kC0 + c0], sh1 = shift[b * kC0 + c0 + 1];
    __syncthreads();
    const int nt = min(kTCH, L0 - t0);
    const size_t e0 = ((size_t)b * L0 + t0) * kC0 + c0;  // element index of (frame t0, channel c0)
    for (int tl = 0; tl < nt; tl += 4) {
        float z0[4], z1[4];
        conv0_quad(xs, tl, w0, w1, z0, z1);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (tl + q < nt) {
                const float y0 = gelu_erf(fmaf(z0[q], sc0, sh0));
                const float y1 = gelu_erf(fmaf(z1[q], sc1, sh1));
                store_planes2(op, y0, y1, e0 + (size_t)(tl + q) * kC0);
            }
        }
    }
}

// "layer_norm" mode: z (+bias) -> LayerNorm over the 512 channels of each frame -> GELU
// (wav2vec2_model.py:2887-2897; Fp32LayerNorm eps 1e-5). Block-wide two-pass reduction per frame quad.
__global__ void __launch_bounds__(256) conv0_apply_ln_kernel(const float* __restrict__ x, long long L, int L0,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ cbias,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const OutPlanes op) {
    __shared__ __align__(16) float xs[kXTile];
    __shared__ float red[8][4];
    __shared__ float stat[4];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kTCH;
    conv0_load_tile(x + (size_t)b * L, L, (long long)t0 * kC0S, xs);
    float w0[kC0K], w1[kC0K];
    const int c0 = 2 * threadIdx.x;
#pragma unroll
    for (int j = 0; j < kC0K; ++j) w0[j] = w[c0 * kC0K + j], w1[j] = w[(c0 + 1) * kC0K + j];
    const float cb0 = cbias ? cbias[c0] : 0.f, cb1 = cbias ? cbias[c0 + 1] : 0.f;
    const float g0 = gamma[c0], g1 = gamma[c0 + 1], be0 = beta[c0], be1 = beta[c0 + 1];
    __syncthreads();
    const int nt = min(kTCH, L0 - t0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t e0 = ((size_t)b * L0 + t0) * kC0 + c0;
    for (int tl = 0; tl < nt; tl += 4) {
        float z0[4], z1[4], mean[4], rstd[4];
        conv0_quad(xs, tl, w0, w1, z0, z1);
#pragma unroll
        for (int q = 0; q < 4; ++q) z0[q] += cb0, z1[q] += cb1;
        // mean over 512 channels
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float s = warp_sum(z0[q] + z1[q]);
            if (lane == 0) red[warp][q] = s;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
            stat[threadIdx.x] = s * (1.0f / kC0);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) mean[q] = stat[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float d0 = z0[q] - mean[q], d1 = z1[q] - mean[q];
            const float s = warp_sum(d0 * d0 + d1 * d1);
            if (lane == 0) red[warp][q] = s;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
            stat[threadIdx.x] = rsqrtf(s * (1.0f / kC0) + 1e-5f);
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) rstd[q] = stat[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (tl + q < nt) {
                const float y0 = gelu_erf(fmaf((z0[q] - mean[q]) * rstd[q], g0, be0));
                const float y1 = gelu_erf(fmaf((z1[q] - mean[q]) * rstd[q], g1, be1));
                store_planes2(op, y0, y1, e0 + (size_t)(tl + q) * kC0);
            }
        }
    }
}

cudaError_t launch_conv0_groupnorm(const float* x, int B, long long L, int L0, const float* w, const float* gamma,
                                   const float* beta, float* ws_part /*[B][nchunk][65]*/,
                                   float* ws_scale_shift /*[2][B][512]*/, const OutPlanes& op, cudaStream_t s) {
    const int nchunk = (L0 + kTCH - 1) / kTCH;
    float* scale = ws_scale_shift;
    float* shift = ws_scale_shift + (size_t)B * kC0;
    dim3 grid(nchunk, B);
    conv0_moments_kernel<<<grid, 256, 0, s>>>(x, L, L0, ws_part, nchunk);
    conv0_finalize_kernel<<<B, kC0, 0, s>>>(ws_part, nchunk, L0, w, gamma, beta, scale, shift);
    conv0_apply_gn_kernel<<<grid, 256, 0, s>>>(x, L, L0, w, scale, shift, op);
    return cudaGetLastError();
}

cudaError_t launch_conv0_layernorm(const float* x, int B, long long L, int L0, const float* w, const float* cbias,
                                   const float* gamma, const float* beta, const OutPlanes& op, cudaStream_t s) {
    const int nchunk = (L0 + kTCH - 1) / kTCH;
    dim3 grid(nchunk, B);
    conv0_apply_ln_kernel<<<grid, 256, 0, s>>>(x, L, L0, w, cbias, gamma, beta, op);
    return cudaGetLastError();
}

size_t conv0_ws_part_floats(int B, int L0) { return (size_t)B * ((L0 + kTCH - 1) / kTCH) * kNMom; }

}  // namespace s3b
