/*
 * s3prl_b200 — C ABI of the Blackwell (sm_100a) upstream feature-extraction hot path.
 *
 * Plain pointers and sizes only (no torch types). Every entry point returns 0 on success and a
 * non-zero status otherwise; s3b_last_error() returns a thread-local message for the last failure.
 * The Python binding a reference maintainer would add is the ctypes stub in INTEGRATION.md
 * (ours: s3prl_b200/lib.py). Citations are to the reference tree s3prl/s3prl @ 0.4.18.
 */
#ifndef S3PRL_B200_H_
#define S3PRL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct s3b_model s3b_model;

/* Architecture of one wav2vec2 / HuBERT / WavLM upstream.
 * Mirrors the fields of HubertConfig (s3prl/upstream/hubert/hubert_model.py:76-278),
 * Wav2Vec2Config (s3prl/upstream/wav2vec2/wav2vec2_model.py:2103-2350) and
 * WavLMConfig (s3prl/upstream/wavlm/WavLM.py:162-245) that the extraction forward reads. */
typedef struct s3b_config {
    int32_t family;               /* 0 = hubert, 1 = wav2vec2 and data2vec, 2 = wavlm, 3 = distiller (selects the frame-mask rule) */
    int32_t extractor_layer_norm; /* 0: extractor_mode "default" (GroupNorm after conv 0); 1: "layer_norm" */
    int32_t conv_bias;            /* conv_bias */
    int32_t layer_norm_first;     /* pre-LN transformer (large_ll60k, wavlm_large) */
    int32_t normalize_wav;        /* task_cfg.normalize: per-utterance F.layer_norm of the waveform */
    int32_t num_layers;           /* encoder_layers */
    int32_t embed_dim;            /* encoder_embed_dim (multiple of 128) */
    int32_t ffn_dim;              /* encoder_ffn_embed_dim */
    int32_t num_heads;            /* encoder_attention_heads; head dim must be 64 */
    int32_t pos_conv_kernel;      /* conv_pos (128; data2vec 95) */
    int32_t pos_conv_groups;      /* conv_pos_groups (16) */
    int32_t relative_position;    /* WavLM relative_position_embedding */
    int32_t num_buckets;          /* WavLM num_buckets (320) */
    int32_t max_distance;         /* WavLM max_distance (800) */
    int32_t gru_rel_pos;          /* WavLM gru_rel_pos */
    /* Distiller / DistilHuBERT (s3prl/upstream/distiller/model.py:81-269): */
    int32_t no_feature_layer_norm; /* 1: post_extract_proj takes the conv features directly (no LayerNorm(512)) */
    int32_t pred_heads;            /* n_tasks prediction heads Linear -> GELU -> SplitLinear on the encoder output (0: none) */
    int32_t pos_conv_depth;        /* 0 / 1: one weight-normed conv of pos_conv_kernel taps (make_conv_pos); > 1 (data2vec):
                                      that many blocks Conv1d(k) -> LayerNorm(no affine) -> GELU with
                                      k = max(3, pos_conv_kernel / pos_conv_depth) (wav2vec2_model.py:2995-3026) */
    int32_t reserved[5];
} s3b_config;

/* Library / error ------------------------------------------------------------------------------ */
int s3b_version(void);
const char* s3b_last_error(void);
/* number of CUDA devices visible (0 on a GPU-less host; never fails) */
int s3b_device_count(void);

/* Model lifetime --------------------------------------------------------------------------------
 * Replaces load_converted_model + HubertModel(...).load_state_dict
 * (s3prl/upstream/hubert/convert.py:37-56, wav2vec2/convert.py:26-39, wavlm/expert.py:37-44).
 * Tensors are passed by their reference state-dict key, fp32, host memory, C-contiguous.
 * Unknown keys (pre-training heads: mask_emb, final_proj, quantizer.*, ...) are ignored.
 * s3b_model_finalize folds weight_norm of pos_conv, re-lays conv weights for channels-last implicit GEMM,
 * splits every GEMM weight into bf16 hi/lo, uploads to the current CUDA device. */
int s3b_model_create(const s3b_config* cfg, s3b_model** out);
int s3b_model_set_tensor(s3b_model* m, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int s3b_model_finalize(s3b_model* m);
void s3b_model_destroy(s3b_model* m);

/* Frame bookkeeping (bit-exact integer rules) ------------------------------------
 * T = conv-stack output length for a padded batch of max length max_len
 * (ConvFeatureExtractionModel, wav2vec2_model.py:2857-2934). Returns -1 if max_len is too short. */
int64_t s3b_num_frames(const s3b_model* m, int64_t max_len);
/* valid_frames[b] = number of un-padded frames of utterance b inside a batch padded to max_len:
 * HuBERT/WavLM chunk-all rule (hubert_model.py:454-464, WavLM.py:339-349) or the wav2vec2 conv-length
 * rule (wav2vec2_model.py:2610-2625,2652-2671). frame t of utterance b is padding iff t >= valid_frames[b]. */
int s3b_valid_frames(const s3b_model* m, const int64_t* lens, int32_t batch, int64_t max_len, int32_t* valid_frames);

/* The hot path ------------------------------------------------------------------------------------
 * Replaces UpstreamExpert.forward + hooks (hubert/expert.py:36-72, wav2vec2/expert.py:61-97,
 * wavlm/expert.py:45-87, interfaces.py:100-131).
 *  wavs        : host array of `batch` DEVICE pointers to fp32 waveforms (un-padded, any order)
 *  lens        : host array of `batch` lengths (samples)
 *  max_len     : padded batch length Lmax (>= max(lens); equal to it for single-GPU use; the global max of
 *                the un-sharded batch when the batch is sharded across ranks)
 *  hidden_out  : DEVICE buffer [s3b_num_outputs(m)][batch][T][embed_dim] fp32, T = s3b_num_frames(max_len):
 *                num_layers+1 hidden states (input of every layer + encoder output); for a distiller model
 *                (distiller/expert.py:44-63) feat_final, the num_layers layer outputs, then the pred_heads predictions
 *  stream      : cudaStream_t the work is enqueued on (asynchronous with respect to the host) */
int32_t s3b_num_outputs(const s3b_model* m);
int s3b_forward(s3b_model* m, const float* const* wavs, const int64_t* lens, int32_t batch, int64_t max_len,
                float* hidden_out, void* stream);
/* Same, end-to-end from HOST buffers: waveforms are host pointers, hidden_out is a host buffer; the
 * host<->device copies are part of the call, which returns after the result has landed. */
int s3b_forward_host(s3b_model* m, const float* const* wavs, const int64_t* lens, int32_t batch, int64_t max_len,
                     float* hidden_out);

/* Options of s3b_forward_ex (zero-initialise, set struct_size = sizeof(s3b_forward_opts)).
 *  lanes            : 0 = default (two utterance micro-batches enqueued alternately on two streams so that one
 *                     micro-batch's small kernels and tails fill the SMs the other leaves idle; bit-identical to
 *                     lanes = 1 because utterances are independent given max_len), 1 or 2
 *  layer_stride     : elements between consecutive hidden states in hidden_out; 0 = batch * T * embed_dim. A caller
 *                     that runs a sub-batch into its slice of a larger [NL+1][Btotal][T][D] buffer passes Btotal*T*D.
 *  ffn_out          : optional DEVICE [num_layers][batch][T][embed_dim]: fc2 output (+bias) of every layer BEFORE
 *                     the residual add = layer_results[i][2] of the reference (wav2vec2/expert.py:87-93,
 *                     feature_selection "fairseq_layers_before_residual"); ffn_layer_stride like layer_stride
 *  last_residual    : optional DEVICE [batch][T][embed_dim], layer_norm_first models only: output of the last layer
 *                     before encoder.layer_norm = layer_results[-1][0] ("fairseq_layers", wav2vec2/expert.py:81-86) */
typedef struct s3b_forward_opts {
    int32_t struct_size;
    int32_t lanes;
    int64_t layer_stride;
    float* ffn_out;
    int64_t ffn_layer_stride;
    float* last_residual;
    int64_t reserved[4];
} s3b_forward_opts;
int s3b_forward_ex(s3b_model* m, const float* const* wavs, const int64_t* lens, int32_t batch, int64_t max_len,
                   float* hidden_out, void* stream, const s3b_forward_opts* opts);
/* Lanes the library picks for a batch when opts->lanes == 0 (two from a measured frame count on, else one;
 * S3B_LANES / S3B_LANE_MIN_FRAMES override). No reference counterpart: scheduling only, results are bit-identical. */
int32_t s3b_default_lanes(const s3b_model* m, int32_t batch, int64_t max_len);
/* s3b_forward_host that also leaves the hidden states in a caller-owned DEVICE buffer [NL+1][batch][T][D]
 * (hidden_out_dev may be NULL = internal staging), so that a device-side consumer (Featurizer, all-gather) can run
 * without re-uploading them. */
int s3b_forward_host_ex(s3b_model* m, const float* const* wavs, const int64_t* lens, int32_t batch, int64_t max_len,
                        float* hidden_out, float* hidden_out_dev);

/* WavLM relative-position bucket of rel[i] = key - query, bidirectional (integer rule, bit-exact with
 * MultiheadAttention._relative_positions_bucket, s3prl/upstream/wavlm/modules.py:418-448). Host arithmetic. */
int s3b_wavlm_buckets(int32_t num_buckets, int32_t max_distance, const int32_t* rel, int32_t n, int32_t* out);

/* Accounting --------------------------------------------------------------------------------------
 * Kernel categories: 0 = tcgen05 GEMM, 1 = tcgen05 attention, 2 = conv-0 (+GroupNorm/LayerNorm+GELU),
 * 3 = LayerNorm kernels, 4 = misc (waveform packing, WavLM gate). With profiling enabled every launch of
 * s3b_forward is bracketed by CUDA events on the launching stream; s3b_profile_read synchronises the device and
 * returns, per category, accumulated device milliseconds, ALGORITHMIC flops (2*M*N*K of the un-padded
 * contraction; 4*T^2*D per utterance and layer for attention) and kernel launches. Arrays have 5 entries. */
int s3b_profile_enable(s3b_model* m, int32_t enable);
int s3b_profile_read(s3b_model* m, double* ms, double* flops, int64_t* launches, int32_t reset);
/* kernels launched by s3b_forward / s3b_forward_host on this model since creation */
int64_t s3b_launch_count(const s3b_model* m);

/* Featurizer (s3prl/upstream/interfaces.py:217-248) --------------------------------------------------
 * out[i] = sum_l w[l] * hs[l][i], hs = [num][n_per_layer] fp32 device, w = device (already softmaxed). */
int s3b_weighted_sum(const float* hs, int32_t num, int64_t n_per_layer, const float* w, float* out, void* stream);
/* grad_w[l] = sum_i hs[l][i] * grad_out[i] */
int s3b_weighted_sum_backward(const float* hs, int32_t num, int64_t n_per_layer, const float* grad_out,
                              float* grad_w, void* stream);

/* Fused Featurizer + all-gather over NVLink peer memory (multi-GPU step) ---------------------------
 * One process per GPU on one NVSwitch node. Every rank creates an exchange object (the library cudaMalloc's
 * [slots][world][block_elems] floats + flags and exports them as a 64-byte CUDA IPC handle), the host side exchanges
 * the handles (torch.distributed.all_gather in s3prl_b200/parallel.py) and calls s3b_peer_connect with all of them.
 *  s3b_peer_push(step): ONE kernel computes sum_l w[l] * hs[l][i] over this rank's block (hs layers layer_stride
 *    elements apart) and stores it into slot step % slots of EVERY rank's buffer at this rank's position, then
 *    releases a per-(slot, rank) flag = step + 1 at system scope.
 *  s3b_peer_wait(step): stream-ordered wait until every rank's flag of that slot has reached step + 1: the local
 *    gathered buffer s3b_peer_slot(step) = [world][block_elems] is then complete.
 * Contract: a slot is rewritten `slots` steps later; the host protocol (parallel.FeatureGatherer) waits for step s-1
 * before pushing step s, so a writer never overtakes a reader by more than slots - 2 steps. */
typedef struct s3b_peer s3b_peer;
int s3b_peer_create(int32_t rank, int32_t world, int64_t block_elems, int32_t slots, s3b_peer** out,
                    void* ipc_handle_out /* 64 bytes */);
int s3b_peer_connect(s3b_peer* p, const void* all_handles /* world x 64 bytes, rank-major */);
float* s3b_peer_slot(s3b_peer* p, uint32_t step);
int s3b_peer_push(s3b_peer* p, const float* hs, int32_t num, int64_t layer_stride, const float* w, uint32_t step,
                  void* stream);
int s3b_peer_wait(s3b_peer* p, uint32_t step, void* stream);
void s3b_peer_destroy(s3b_peer* p);

/* fbank baseline upstream (s3prl/upstream/baseline/expert.py:69-79 over torchaudio.compliance.kaldi.fbank,
 * fbank.yaml: 80 mel bins, 25 ms / 10 ms, log; + 2 x ComputeDeltas(5) + per-utterance CMVN).
 *  wavs : host array of `batch` DEVICE pointers (fp32, un-padded), lens : host array of lengths
 *  out  : DEVICE buffer [batch][s3b_fbank_num_frames(max(lens))][240] fp32, zero beyond each utterance's frames */
int64_t s3b_fbank_num_frames(int64_t len);
int s3b_fbank(const float* const* wavs, const int64_t* lens, int32_t batch, float* out, void* stream);

/* mel / linear baseline upstreams (s3prl/upstream/baseline/expert.py:52-79 over OnlinePreprocessor.forward,
 * s3prl/upstream/baseline/preprocessor.py:150-223; mel.yaml / linear.yaml). The host side (Python, like the
 * reference) derives the data-dependent lengths; these entry points do the arithmetic.
 *  s3b_trimmed_lengths: out[b] = (index of the last non-zero sample) + 1, or lens[b] if all zero (synchronous).
 *  s3b_melspec: STFT power (n_fft 400, hop 160, hann, center/reflect over the row padded to `padded_len`) ->
 *    mel != 0: 80-bin HTK mel, else 201 linear bins -> log(x + 1e-10) -> CMVN over the first feats_len[b] frames
 *    -> first final_len[b] frames kept, zero up to t_out. out: DEVICE [batch][t_out][80 | 201] fp32. */
int s3b_trimmed_lengths(const float* const* wavs, const int64_t* lens, int32_t batch, int64_t* out);
int s3b_melspec(const float* const* wavs, const int64_t* trimmed_lens, int32_t batch, int64_t padded_len, int32_t mel,
                const int32_t* feats_len, const int32_t* final_len, int32_t t_out, float* out, void* stream);

/* Building blocks exposed for parity tests (device pointers, fp32) -------------------------------------- */
/* out[M][N] = act(A[M][K] * W[N][K]^T + bias[N]) (+ residual[M][N]); K % 64 == 0, N % 16 == 0 */
int s3b_linear_f32(const float* a, const float* w, const float* bias, const float* residual, int64_t m, int32_t n,
                   int32_t k, int32_t gelu, float* out, void* stream);
/* micro-benchmark of the production GEMM kernel: `iters` back-to-back launches of [M][K] x [N][K]^T with the fc1-style
 * epilogue (gelu != 0: bias + GELU + bf16 hi/lo output) or the out_proj-style one (bias + residual + fp32 output).
 * force_un in {0 = the library's choice, 128, 256} forces the CTA-pair tile width. out[0] = ms per launch,
 * out[1] = tile width used. (tools/gemm_tile_sweep.py) */
int s3b_gemm_bench(int64_t m, int32_t n, int32_t k, int32_t gelu, int32_t force_un, int32_t iters, float* out);
/* y = LayerNorm_D(x) (eps 1e-5), optional GELU; D in {512,768,1024,1280} */
int s3b_layernorm_f32(const float* x, int64_t m, int32_t d, const float* gamma, const float* beta, int32_t gelu,
                      float* out, void* stream);
/* softmax(q k^T / 8 + key padding mask) v per (batch, head); q,k,v: [batch][T][heads*64] fp32 */
int s3b_attention_f32(const float* q, const float* k, const float* v, const int32_t* valid_frames, int32_t batch,
                      int32_t t, int32_t heads, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S3PRL_B200_H_ */
