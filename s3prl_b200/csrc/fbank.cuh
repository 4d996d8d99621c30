// fbank.cu launcher
#pragma once
#include <cuda_runtime.h>

namespace s3b {
// out: [B][max_frames][240] fp32 (80 log-mel | 80 delta | 80 delta-delta, CMVN applied, zero beyond each length)
cudaError_t launch_fbank(const float* const* wavs_dev, const long long* lens_dev, int B, int max_frames, float* out,
                         cudaStream_t s);
}  // namespace s3b

namespace s3b {
// melspec.cu
cudaError_t launch_trimmed_lengths(const float* const* wavs_dev, const long long* lens_dev, int B, long long* out_dev,
                                   cudaStream_t s);
// tmp: [B][1 + Lp/160][dim] scratch; out: [B][t_out][dim]; dim = 80 (mel != 0) or 201
cudaError_t launch_melspec(const float* const* wavs_dev, const long long* trim_lens_dev, int B, long long Lp, int mel,
                           const int* feats_len_dev, const int* final_len_dev, int t_out, float* tmp, float* out,
                           cudaStream_t s);
}  // namespace s3b
