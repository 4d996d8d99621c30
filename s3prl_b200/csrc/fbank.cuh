// fbank.cu launcher
#pragma once
#include <cuda_runtime.h>

namespace s3b {
// out: [B][max_frames][240] fp32 (80 log-mel | 80 delta | 80 delta-delta, CMVN applied, zero beyond each length)
cudaError_t launch_fbank(const float* const* wavs_dev, const long long* lens_dev, int B, int max_frames, float* out,
                         cudaStream_t s);
}  // namespace s3b
