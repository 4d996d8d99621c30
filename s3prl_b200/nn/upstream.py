"""Padded-tensor interface to the B200 upstreams.

Mirror of the reference's new-style wrapper ``s3prl.nn.S3PRLUpstream`` (s3prl/nn/upstream.py:38-231): same
constructor arguments, ``num_layers`` / ``hidden_sizes`` / ``downsample_rates`` properties and the same length
bookkeeping, resolved against ``s3prl_b200.hub`` instead of ``s3prl.hub``; plus its ``Featurizer`` (:234-349, layer
selection + optional per-layer layer_norm + trainable softmax-weighted sum, here through the fused
``s3b_weighted_sum`` kernel and its backward) and ``UpstreamDownstreamModel`` (:352-384):

    model = S3PRLUpstream("hubert_base").cuda()
    all_hs, all_lens = model(wavs_padded, wavs_len)     # wavs_padded [B, Lmax] fp32 CUDA, wavs_len [B] long
    hs, hs_len = Featurizer(model).cuda()(all_hs, all_lens)

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
            # entries of hidden_states: NL + 1, or feat_final + layer outputs + prediction heads for the Distiller
            arch = getattr(self.upstream, "arch", None)
            self._num_layers = arch.num_outputs if arch is not None else self.upstream.num_layers + 1
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


class Featurizer(nn.Module):
    """``s3prl.nn.Featurizer`` (s3prl/nn/upstream.py:234-349): reduce the upstream's layers to one sequence. One layer
    is passed through; several get a trainable softmax-weighted sum over ``layer_selections`` (all layers by default),
    optionally after a per-layer ``F.layer_norm`` over the hidden dimension. The sum streams the layers once through
    the fused CUDA kernel (``s3b_weighted_sum``; gradients for the weights and the features in its backward)."""

    def __init__(self, upstream: S3PRLUpstream, layer_selections: List[int] = None, normalize: bool = False):
        super().__init__()
        assert len(set(upstream.hidden_sizes)) == 1
        assert len(set(upstream.downsample_rates)) == 1
        self._output_size = upstream.hidden_sizes[0]
        self._downsample_rate = upstream.downsample_rates[0]
        self.normalize = normalize
        if upstream.num_layers > 1:
            if layer_selections is not None:
                assert upstream.num_layers >= len(layer_selections)
                self.layer_selections = sorted(layer_selections)
            else:
                self.layer_selections = list(range(upstream.num_layers))
            self.weights = nn.Parameter(torch.zeros(len(self.layer_selections)))

    @property
    def output_size(self) -> int:
        return self._output_size

    @property
    def downsample_rate(self) -> int:
        return self._downsample_rate

    def _weighted_sum(self, all_hs: List[torch.Tensor], all_lens: List[torch.Tensor]):
        from ..upstream.featurizer import weighted_sum

        assert len(all_hs) == len(all_lens) > 1
        if self.normalize:  # the reference normalises the stacked tensor over its last dimension (:319-320)
            all_hs = [F.layer_norm(h, (h.shape[-1],)) for h in all_hs]
        return weighted_sum(all_hs, F.softmax(self.weights, dim=-1)), all_lens[0]

    def forward(self, all_hs: List[torch.Tensor], all_lens: List[torch.Tensor]):
        if len(all_hs) == 1:
            return all_hs[0], all_lens[0]
        all_hs = [h for idx, h in enumerate(all_hs) if idx in self.layer_selections]
        all_lens = [l for idx, l in enumerate(all_lens) if idx in self.layer_selections]
        return self._weighted_sum(all_hs, all_lens)


class UpstreamDownstreamModel(nn.Module):
    """``s3prl.nn.UpstreamDownstreamModel`` (s3prl/nn/upstream.py:352-384). The upstreams of this package are frozen
    (no autograd through the CUDA forward): ``upstream_trainable=True`` is refused instead of training silently on
    constants."""

    def __init__(self, upstream: S3PRLUpstream, featurizer: Featurizer, downstream, upstream_trainable: bool = False):
        super().__init__()
        if upstream_trainable:
            raise NotImplementedError("s3prl_b200 upstreams are frozen: upstream_trainable=True is not supported")
        self.upstream = upstream
        self.featurizer = featurizer
        self.downstream = downstream
        self.upstream_trainable = False

    @property
    def input_size(self):
        return 1

    @property
    def downsample_rate(self):
        return self.featurizer.downsample_rate

    @property
    def output_size(self):
        return self.downstream.output_size

    def forward(self, wav, wav_len, *args, **kwargs):
        with torch.no_grad():
            self.upstream.eval()
            hs, hs_len = self.upstream(wav, wav_len)
        h, h_len = self.featurizer(hs, hs_len)
        return self.downstream(h, h_len, *args, **kwargs)
