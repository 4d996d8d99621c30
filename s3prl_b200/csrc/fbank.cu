// Kaldi-compatible log-mel filterbank + deltas + CMVN (the `fbank` baseline upstream).
// Replaces the per-utterance Python loop of s3prl/upstream/baseline/expert.py:46-50,69-79 over
//   torchaudio.compliance.kaldi.fbank(num_mel_bins=80, frame_length=25, frame_shift=10, use_log_fbank=True)
//     (torchaudio/compliance/kaldi.py: _get_window :154-218, get_mel_banks :436-512, fbank :514-646)
//   2 x ComputeDeltas(win_length=5), concatenated           s3prl/upstream/baseline/extracter.py:58-76
//   per-utterance CMVN with unbiased std                    s3prl/upstream/baseline/extracter.py:78-90
// One warp computes one 25 ms frame: DC removal, pre-emphasis, povey window, 512-point radix-2 FFT in shared
// memory, power spectrum, sparse triangular mel projection (the "matmul"), log. All fp32.
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fbank.cuh"

namespace s3b {

static constexpr int kWin = 400, kShift = 160, kFft = 512, kBins = 257, kMel = 80, kFeat = 240;
static constexpr int kMaxMelNnz = 1024;

struct FbankTables {
    float window[kWin];
    float tw_re[kFft / 2], tw_im[kFft / 2];  // exp(-2 pi i k / 512)
    int mel_start[kMel], mel_count[kMel], mel_off[kMel];
    float mel_w[kMaxMelNnz];
};
__device__ FbankTables g_tab;
static bool g_tab_ready[64] = {false};

static void build_tables(FbankTables& t) {
    // povey window: hann(400, periodic=False) ** 0.85   (kaldi.py:98-100)
    for (int n = 0; n < kWin; ++n) {
        const double h = 0.5 - 0.5 * cos(2.0 * M_PI * n / (kWin - 1));
        t.window[n] = (float)pow(h, 0.85);
    }
    for (int k = 0; k < kFft / 2; ++k) {
        t.tw_re[k] = (float)cos(-2.0 * M_PI * k / kFft);
        t.tw_im[k] = (float)sin(-2.0 * M_PI * k / kFft);
    }
    // mel banks (kaldi.py:436-512): low 20 Hz, high = nyquist 8000 Hz, 80 bins over 256 fft bins, fp32 like torch
    const float low = 20.0f, high = 8000.0f, fft_bin_width = 16000.0f / kFft;
    auto mel = [](float f) { return 1127.0f * logf(1.0f + f / 700.0f); };
    const double mel_low = 1127.0 * log(1.0 + (double)low / 700.0), mel_high = 1127.0 * log(1.0 + (double)high / 700.0);
    const double delta = (mel_high - mel_low) / (kMel + 1);
    int off = 0;
    for (int j = 0; j < kMel; ++j) {
        const float left = (float)(mel_low + j * delta), center = (float)(mel_low + (j + 1.0) * delta),
                    right = (float)(mel_low + (j + 2.0) * delta);
        int start = -1, count = 0;
        for (int i = 0; i < kFft / 2; ++i) {
            const float m = mel(fft_bin_width * (float)i);
            const float up = (m - left) / (center - left), down = (right - m) / (right - center);
            const float w = fmaxf(0.0f, fminf(up, down));
            if (w > 0.0f) {
                if (start < 0) start = i;
                // weights of one triangle are contiguous in i
                if (off + count < kMaxMelNnz) t.mel_w[off + count] = w;
                ++count;
            }
        }
        t.mel_start[j] = start < 0 ? 0 : start, t.mel_count[j] = count, t.mel_off[j] = off;
        off += count;
    }
}

static cudaError_t ensure_tables(cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_tab_ready[dev]) return cudaSuccess;
    static FbankTables host;
    build_tables(host);
    e = cudaMemcpyToSymbolAsync(g_tab, &host, sizeof(host), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && dev < 64) g_tab_ready[dev] = true;
    return e;
}

__device__ __forceinline__ int bitrev9(int x) { return (int)(__brev((unsigned)x) >> 23); }

// grid: (ceil(max_frames / 4), B); block: 128 threads = 4 warps = 4 frames
__global__ void __launch_bounds__(128) fbank_frames_kernel(const float* const* __restrict__ wavs,
                                                           const long long* __restrict__ lens, int max_frames,
                                                           float* __restrict__ out) {
    __shared__ float s_re[4][kFft], s_im[4][kFft], s_pw[4][kBins + 3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int t = blockIdx.x * 4 + warp;
    const long long L = lens[b];
    const int m = L < kWin ? 0 : (int)(1 + (L - kWin) / kShift);  // snip_edges=True (kaldi.py:_get_strided)
    if (t >= m) return;  // warp-uniform
    const float* w = wavs[b] + (size_t)t * kShift;
    float* re = s_re[warp];
    float* im = s_im[warp];

    // frame -> registers (13 samples per lane, last partial), DC removal
    float x[13];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int n = lane + 32 * i;
        x[i] = n < kWin ? w[n] : 0.f;
        sum += x[i];
    }
    const float mean = warp_sum(sum) / (float)kWin;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int n = lane + 32 * i;
        if (n < kWin) re[n] = x[i] - mean;
    }
    __syncwarp();
    // pre-emphasis (replicate first sample), povey window, zero pad, bit-reversed placement for the DIT FFT
    float y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = lane + 32 * i;
        float v = 0.f;
        if (n < kWin) v = (re[n] - 0.97f * re[n > 0 ? n - 1 : 0]) * g_tab.window[n];
        y[i] = v;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int n = lane + 32 * i;
        const int r = bitrev9(n);
        re[r] = y[i];
        im[r] = 0.f;
    }
    __syncwarp();
    // 9 radix-2 stages, 256 butterflies each (8 per lane)
    for (int s = 1; s <= 9; ++s) {
        const int half = 1 << (s - 1);
        const int tw_stride = kFft >> s;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int bf = lane + 32 * i;            // butterfly index 0..255
            const int j = bf & (half - 1);
            const int base = ((bf >> (s - 1)) << s) + j;
            const float wr = g_tab.tw_re[j * tw_stride], wi = g_tab.tw_im[j * tw_stride];
            const float ar = re[base], ai = im[base];
            const float br = re[base + half], bi = im[base + half];
            const float tr = br * wr - bi * wi, ti = br * wi + bi * wr;
            re[base] = ar + tr, im[base] = ai + ti;
            re[base + half] = ar - tr, im[base + half] = ai - ti;
        }
        __syncwarp();
    }
    // power spectrum (use_power=True): |X|^2 ; kaldi.py computes abs() then pow(2)
    float* pw = s_pw[warp];
    for (int k = lane; k < kBins; k += 32) {
        const float a = sqrtf(re[k] * re[k] + im[k] * im[k]);
        pw[k] = a * a;
    }
    __syncwarp();
    // mel projection + log(max(e, eps))
    float* o = out + ((size_t)b * max_frames + t) * kFeat;
    for (int j = lane; j < kMel; j += 32) {
        const int st = g_tab.mel_start[j], cnt = g_tab.mel_count[j];
        const float* mw = g_tab.mel_w + g_tab.mel_off[j];
        float acc = 0.f;
        for (int i = 0; i < cnt; ++i) acc = fmaf(pw[st + i], mw[i], acc);
        o[j] = logf(fmaxf(acc, 1.1920928955078125e-07f));
    }
}

// ComputeDeltas(win_length=5, mode="replicate") applied to columns [src, src+80) -> [src+80, src+160)
__global__ void delta_kernel(float* __restrict__ feat, const long long* __restrict__ lens, int max_frames, int src) {
    const int b = blockIdx.y;
    const long long L = lens[b];
    const int m = L < kWin ? 0 : (int)(1 + (L - kWin) / kShift);
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = idx / kMel, j = idx - t * kMel;
    if (t >= m) return;
    const float* c = feat + (size_t)b * max_frames * kFeat + src + j;
    auto at = [&](int tt) { return c[(size_t)min(max(tt, 0), m - 1) * kFeat]; };
    // F.conv1d with kernel [-2,-1,0,1,2] then / 10  (torchaudio functional.compute_deltas)
    const float v = (-2.f * at(t - 2) - at(t - 1) + 0.f * at(t) + at(t + 1) + 2.f * at(t + 2)) / 10.0f;
    feat[((size_t)b * max_frames + t) * kFeat + src + kMel + j] = v;
}

// CMVN per utterance and feature: (x - mean) / (1e-10 + std_unbiased); frames >= m are zero (pad_sequence)
__global__ void __launch_bounds__(256) cmvn_kernel(float* __restrict__ feat, const long long* __restrict__ lens,
                                                   int max_frames) {
    const int b = blockIdx.x;
    const int j = threadIdx.x;
    const long long L = lens[b];
    const int m = L < kWin ? 0 : (int)(1 + (L - kWin) / kShift);
    if (j >= kFeat) return;
    float* c = feat + (size_t)b * max_frames * kFeat + j;
    double s = 0.0;
    for (int t = 0; t < m; ++t) s += (double)c[(size_t)t * kFeat];
    const float mean = m > 0 ? (float)(s / m) : 0.f;
    double q = 0.0;
    for (int t = 0; t < m; ++t) {
        const float d = c[(size_t)t * kFeat] - mean;
        q += (double)(d * d);
    }
    const float sd = (float)sqrt(q / (double)(m - 1));  // m == 1 -> NaN, as torch.std does
    const float inv = 1.0f / (1e-10f + sd);
    for (int t = 0; t < m; ++t) c[(size_t)t * kFeat] = (c[(size_t)t * kFeat] - mean) * inv;
    for (int t = m; t < max_frames; ++t) c[(size_t)t * kFeat] = 0.f;
}

cudaError_t launch_fbank(const float* const* wavs_dev, const long long* lens_dev, int B, int max_frames, float* out,
                         cudaStream_t s) {
    cudaError_t e = ensure_tables(s);
    if (e != cudaSuccess) return e;
    if (max_frames <= 0 || B <= 0) return cudaSuccess;
    dim3 g1((max_frames + 3) / 4, B);
    fbank_frames_kernel<<<g1, 128, 0, s>>>(wavs_dev, lens_dev, max_frames, out);
    dim3 g2((max_frames * kMel + 255) / 256, B);
    delta_kernel<<<g2, 256, 0, s>>>(out, lens_dev, max_frames, 0);
    delta_kernel<<<g2, 256, 0, s>>>(out, lens_dev, max_frames, kMel);
    cmvn_kernel<<<B, 256, 0, s>>>(out, lens_dev, max_frames);
    return cudaGetLastError();
}

}  // namespace s3b
