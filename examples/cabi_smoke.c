/*
 * Plain-C user of the C ABI (no Python, no torch): proves that include/s3prl_b200.h is valid C and that every entry
 * point links. On a GPU-less host it only queries the library; with a device it would go on to
 * s3b_model_create / set_tensor / finalize / s3b_forward_host (see INTEGRATION.md for the full sequence).
 *
 *   gcc -std=c99 -Wall -Werror -Iinclude examples/cabi_smoke.c -Ls3prl_b200/_lib -ls3prl_b200 \
 *       -Wl,-rpath,$PWD/s3prl_b200/_lib -o cabi_smoke && ./cabi_smoke
 */
#include <stdio.h>
#include <string.h>

#include "s3prl_b200.h"

int main(void) {
    /* take the address of every entry point so that a missing export fails at link time */
    const void* entry[] = {
        (const void*)s3b_version,        (const void*)s3b_last_error,      (const void*)s3b_device_count,
        (const void*)s3b_model_create,   (const void*)s3b_model_set_tensor, (const void*)s3b_model_finalize,
        (const void*)s3b_model_destroy,  (const void*)s3b_num_frames,      (const void*)s3b_valid_frames,
        (const void*)s3b_forward,        (const void*)s3b_forward_host,    (const void*)s3b_profile_enable,
        (const void*)s3b_profile_read,   (const void*)s3b_launch_count,    (const void*)s3b_weighted_sum,
        (const void*)s3b_weighted_sum_backward, (const void*)s3b_fbank,    (const void*)s3b_fbank_num_frames,
        (const void*)s3b_trimmed_lengths, (const void*)s3b_melspec,        (const void*)s3b_linear_f32,
        (const void*)s3b_layernorm_f32,  (const void*)s3b_attention_f32,  (const void*)s3b_forward_ex,
        (const void*)s3b_forward_host_ex, (const void*)s3b_wavlm_buckets, (const void*)s3b_peer_create,
        (const void*)s3b_peer_connect,   (const void*)s3b_peer_slot,       (const void*)s3b_peer_push,
        (const void*)s3b_peer_wait,      (const void*)s3b_peer_destroy,    (const void*)s3b_gemm_bench,
        (const void*)s3b_num_outputs,    (const void*)s3b_default_lanes,
    };
    s3b_config cfg;
    int64_t lens[2] = {16000, 800};
    int32_t valid[2] = {0, 0};
    memset(&cfg, 0, sizeof(cfg));
    printf("s3prl_b200 C ABI version %d, %d entry points, %d CUDA device(s)\n", s3b_version(),
           (int)(sizeof(entry) / sizeof(entry[0])), s3b_device_count());
    /* integer rules need no device: frames of a 1 s utterance; HuBERT frame-mask rule (hubert_model.py:454-464) for a
     * ragged pair, through a model handle that is never finalized (no GPU needed until s3b_model_finalize) */
    if (s3b_num_frames(NULL, 16000) != 49) return 1;
    cfg.family = 0, cfg.num_layers = 12, cfg.embed_dim = 768, cfg.ffn_dim = 3072, cfg.num_heads = 12;
    cfg.pos_conv_kernel = 128, cfg.pos_conv_groups = 16;
    {
        s3b_model* m = NULL;
        if (s3b_model_create(&cfg, &m) != 0) {
            printf("s3b_model_create: %s\n", s3b_last_error());
            return 2;
        }
        if (s3b_valid_frames(m, lens, 2, 16000, valid) != 0) {
            printf("s3b_valid_frames: %s\n", s3b_last_error());
            return 3;
        }
        printf("valid frames of (16000, 800) samples: %d %d\n", (int)valid[0], (int)valid[1]);
        s3b_model_destroy(m);
        if (valid[0] != 49 || valid[1] != 3) return 4; /* 800 samples cover frames 0..2 (chunks of 326 samples) */
    }
    return 0;
}
