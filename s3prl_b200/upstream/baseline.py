"""``fbank`` baseline upstream on the GPU (s3prl/upstream/baseline/expert.py:23-79 with fbank.yaml).

    expert = FbankExpert()
    out = expert([wav_0, wav_1, ...])      # list of 1-D fp32 CUDA tensors
    out["hidden_states"] == [X],  X: [B, max_frames, 240]  (80 log-mel + delta + delta-delta, CMVN)
    expert.get_downsample_rates("hidden_states") == 160
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch
import torch.nn as nn

from .. import lib as _lib


class FbankExpert(nn.Module):
    output_dim = 240
    downsample_rate = 160

    def __init__(self, model_config=None, **kwargs):
        super().__init__()
        self.register_buffer("_device_anchor", torch.zeros(1), persistent=False)

    def get_downsample_rates(self, key: str) -> int:
        return self.downsample_rate

    def forward(self, wavs: List[torch.Tensor]) -> Dict[str, object]:
        lib = _lib.load()
        _lib.require_gpu()
        device = wavs[0].device
        if device.type != "cuda":
            raise _lib.S3BError("s3prl_b200 fbank runs on CUDA devices only (no CPU fallback)")
        wavs = [w.detach().to(torch.float32).contiguous() for w in wavs]
        lens = [int(w.numel()) for w in wavs]
        B = len(wavs)
        m = lib.s3b_fbank_num_frames(max(lens))
        if m < 1:
            raise ValueError("waveforms shorter than one 25 ms frame")
        with torch.cuda.device(device):
            out = torch.empty((B, m, self.output_dim), dtype=torch.float32, device=device)
            ptrs = (C.c_void_p * B)(*[w.data_ptr() for w in wavs])
            lens_c = (C.c_int64 * B)(*lens)
            _lib.check(
                lib.s3b_fbank(ptrs, lens_c, B, C.c_void_p(out.data_ptr()),
                              C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
            )
        return {"last_hidden_state": out, "hidden_states": [out]}
