"""Padded-tensor interface to the B200 upstreams.

Mirror of the reference's new-style wrapper ``s3prl.nn.S3PRLUpstream`` (s3prl/nn/upstream.py:38-231): same
constructor arguments, ``num_layers`` / ``hidden_sizes`` / ``downsample_rates`` properties and the same length
bookkeeping, resolved against ``s3prl_b200.hub`` instead of ``s3prl.hub``:

    model = S3PRLUpstream("hubert_base").cuda()
    all_hs, all_lens = model(wavs_padded, wavs_len)     # wavs_padded [B, Lmax] fp32 CUDA, wavs_len [B] long

Length rules (bit-exact): each layer is trimmed / last-frame-repeated to
``len(range(0, Lmax, stride))`` frames and ``h_len = (len - 1) // stride + 1``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import hub

MIN_SECOND = 0.05
SAMPLE_RATE = 16000


class S3PRLUpstream(nn.Module):
    @classmethod
    def available_names(cls, only_registered_ckpt: bool = False) -> List[str]:
        return hub.options()

    def __init__(
        self,
        name: str,
        path_or_url: str = None,
        refresh: bool = False,
        normalize: bool = False,
        extra_conf: dict = None,
        randomize: bool = False,
    ):
        super().__init__()
        conf = {"refresh": refresh, **(extra_conf or {})}
        if path_or_url is not None:
            conf["ckpt"] = path_or_url
        if randomize:
            # the reference re-initialises the loaded model's parameters (nn/upstream.py:119-122). Our upstreams own
            # their weights: draw a seed from the default generator (reproducible under torch.manual_seed, global
            # seed untouched) and fabricate fresh weights — for the checkpoint's architecture when a path is given.
            conf["randomize_seed"] = int(torch.randint(0, 2**31 - 1, ()).item())
        self.upstream = hub.ENTRIES[name](**conf)
        self.normalize = normalize
        # static facts; the reference discovers them with a pseudo forward (nn/upstream.py:124-128), which would
        # need a GPU at construction time
        if hasattr(self.upstream, "num_layers"):
            self._num_layers = self.upstream.num_layers + 1
            self._hidden_sizes = [self.upstream.hidden_size] * self._num_layers
        else:  # fbank
            self._num_layers = 1
            self._hidden_sizes = [self.upstream.output_dim]
        rate = self.upstream.get_downsample_rates("hidden_states")
        self._downsample_rates = [rate] * self._num_layers

    @property
    def num_layers(self) -> int:
        return self._num_layers

    @property
    def downsample_rates(self) -> List[int]:
        return self._downsample_rates

    @property
    def hidden_sizes(self) -> List[int]:
        return self._hidden_sizes

    @staticmethod
    def _match_length(xs: torch.Tensor, target_max_len: int) -> torch.Tensor:
        xs_max_len = xs.size(1)
        if xs_max_len > target_max_len:
            assert xs_max_len // target_max_len == 1, f"{xs_max_len}, {target_max_len}"
            xs = xs[:, :target_max_len, :]
        elif xs_max_len < target_max_len:
            assert target_max_len // xs_max_len == 1, f"{target_max_len}, {xs_max_len}"
            xs = torch.cat((xs, xs[:, -1:, :].repeat(1, target_max_len - xs_max_len, 1)), dim=1)
        return xs

    def forward(self, wavs: torch.Tensor, wavs_len: torch.Tensor) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        if wavs.dim() == 3:
            wavs = wavs.squeeze(-1)
        original_wavs_len = wavs_len
        if int(max(original_wavs_len)) < MIN_SECOND * SAMPLE_RATE:
            padded = int(MIN_SECOND * SAMPLE_RATE) - int(max(original_wavs_len))
            wavs = torch.cat((wavs, wavs.new_zeros(wavs.size(0), padded)), dim=1)
            wavs_len = wavs_len + padded
        wavs_list = [wav[: int(n)] for wav, n in zip(wavs, wavs_len)]
        hidden_states = self.upstream(wavs_list)["hidden_states"]
        assert len(hidden_states) == self.num_layers, f"{len(hidden_states)}, {self.num_layers}"
        max_wav_len = int(max(wavs_len))
        all_hs, all_lens = [], []
        for h, stride in zip(hidden_states, self.downsample_rates):
            expected = len(range(0, max_wav_len, stride))
            h = self._match_length(h, expected)
            h_len = torch.div(original_wavs_len - 1, stride, rounding_mode="floor") + 1
            h = h[:, : int(max(h_len)), :]
            if self.normalize:
                h = F.layer_norm(h, h.shape[-1:])
            all_hs.append(h)
            all_lens.append(h_len)
        return all_hs, all_lens
