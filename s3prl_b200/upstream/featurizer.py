"""Featurizer: layer selection + softmax-weighted sum of the hidden states + per-utterance un-padding.

Mirror of ``s3prl.upstream.interfaces.Featurizer`` (s3prl/upstream/interfaces.py:134-272) for upstreams of this
package: same constructor arguments, ``forward(paired_wavs, paired_features) -> List[Tensor[T_i, D]]``, same
``weights`` parameter (trainable; initial zeros), same ``round(len / downsample_rate)`` length rule. The
weighted sum streams the NL+1 layers once through ``s3b_weighted_sum`` instead of ``torch.stack`` + mul + sum,
and its backward (gradient of the layer weights and of the features) is provided so SUPERB training
(run_downstream.py, featurizer trainable, upstream frozen) works.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from .. import lib as _lib

SAMPLE_RATE = 16000
TOLERABLE_SEQLEN_DIFF = 5


def _stacked_view(feature: Sequence[Tensor]):
    """Hidden states returned by UpstreamExpert are consecutive slices of one [NL+1, B, T, D] buffer; detect that
    so no copy is needed (otherwise fall back to one torch.stack, still a single extra pass)."""
    f0 = feature[0]
    n = f0.numel()
    ok = all(
        f.is_contiguous() and f.dtype == torch.float32 and f.shape == f0.shape
        and f.data_ptr() == f0.data_ptr() + i * n * 4
        for i, f in enumerate(feature)
    )
    base = getattr(f0, "_base", None)
    if ok and base is not None and base.is_contiguous() and base.data_ptr() == f0.data_ptr() and base.numel() >= n * len(feature):
        return base.reshape(-1)[: n * len(feature)].view(len(feature), *f0.shape)
    return torch.stack([f.to(torch.float32) for f in feature], dim=0).contiguous()


class _WeightedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, stacked: Tensor, norm_weights: Tensor) -> Tensor:
        lib = _lib.load()
        num = stacked.shape[0]
        n = stacked[0].numel()
        out = torch.empty_like(stacked[0])
        w = norm_weights.detach().to(stacked.device, torch.float32).contiguous()
        with torch.cuda.device(stacked.device):
            _lib.check(
                lib.s3b_weighted_sum(
                    C.c_void_p(stacked.data_ptr()), num, n, C.c_void_p(w.data_ptr()), C.c_void_p(out.data_ptr()),
                    C.c_void_p(torch.cuda.current_stream(stacked.device).cuda_stream),
                )
            )
        ctx.save_for_backward(stacked, w)
        ctx.weights_device = norm_weights.device
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        stacked, w = ctx.saved_tensors
        lib = _lib.load()
        num = stacked.shape[0]
        n = stacked[0].numel()
        grad_w = None
        grad_stacked = None
        g = grad_out.to(torch.float32).contiguous()
        if ctx.needs_input_grad[1]:
            grad_w = torch.empty(num, device=stacked.device, dtype=torch.float32)
            with torch.cuda.device(stacked.device):
                _lib.check(
                    lib.s3b_weighted_sum_backward(
                        C.c_void_p(stacked.data_ptr()), num, n, C.c_void_p(g.data_ptr()), C.c_void_p(grad_w.data_ptr()),
                        C.c_void_p(torch.cuda.current_stream(stacked.device).cuda_stream),
                    )
                )
            grad_w = grad_w.to(ctx.weights_device)  # the Featurizer's weights may live on another device
        if ctx.needs_input_grad[0]:
            grad_stacked = w.view(-1, *([1] * g.dim())) * g.unsqueeze(0)
        return grad_stacked, grad_w


def weighted_sum(feature: Sequence[Tensor], norm_weights: Tensor) -> Tensor:
    """sum_l norm_weights[l] * feature[l] on the GPU through the C ABI (n % 4 == 0 required by the kernel)."""
    stacked = _stacked_view(feature)
    if not stacked.is_cuda:
        raise _lib.S3BError("s3prl_b200 Featurizer needs CUDA tensors (no CPU fallback)")
    return _WeightedSum.apply(stacked, norm_weights)


class Featurizer(nn.Module):
    def __init__(
        self,
        upstream: nn.Module,
        feature_selection: str = "hidden_states",
        upstream_device: str = "cuda",
        layer_selection: int = None,
        normalize: bool = False,
        **kwargs,
    ):
        super().__init__()
        self.name = "Featurizer"
        upstream.eval()
        paired_wavs = [torch.randn(SAMPLE_RATE).to(upstream_device)]
        with torch.no_grad():
            paired_features = upstream(paired_wavs)
        if feature_selection not in paired_features:
            if "hidden_states" in paired_features:
                feature_selection = "hidden_states"
            else:
                raise ValueError(f"{feature_selection} is not a supported feature selection")
        self.feature_selection = feature_selection
        self.layer_selection = layer_selection
        self.normalize = normalize
        feature = self._select_feature(paired_features)
        if isinstance(feature, (list, tuple)):
            self.layer_num = len(feature)
            self.weights = nn.Parameter(torch.zeros(self.layer_num))
            feature = self._weighted_sum(list(feature))
        self.output_dim = feature.size(-1)
        if hasattr(upstream, "get_downsample_rates"):
            self.downsample_rate = upstream.get_downsample_rates(feature_selection)
        else:
            self.downsample_rate = round(max(len(wav) for wav in paired_wavs) / feature.size(1))

    def _select_feature(self, features: Dict):
        feature = features.get(self.feature_selection)
        if isinstance(feature, dict):
            feature = list(feature.values())
        if isinstance(feature, (list, tuple)) and len(feature) == 1:
            feature = feature[0]
        if isinstance(feature, (list, tuple)) and isinstance(self.layer_selection, int):
            feature = feature[self.layer_selection]
        return feature

    def _weighted_sum(self, feature: List[Tensor]) -> Tensor:
        assert self.layer_num == len(feature)
        if self.normalize:
            feature = [F.layer_norm(f, (f.shape[-1],)) for f in feature]
        norm_weights = F.softmax(self.weights, dim=-1)
        return weighted_sum(feature, norm_weights)

    def tolist(self, paired_wavs: List[Tensor], paired_feature: Tensor) -> List[Tensor]:
        assert paired_feature.dim() == 3, "(batch_size, max_seq_len, feat_dim)"
        feature_len = [round(len(wav) / self.downsample_rate) for wav in paired_wavs]
        length_diff = abs(
            paired_feature.size(1) - round(max([len(wav) for wav in paired_wavs]) / self.downsample_rate)
        )
        assert length_diff < TOLERABLE_SEQLEN_DIFF, f"{length_diff} >= {TOLERABLE_SEQLEN_DIFF}"
        return [f[:l] for f, l in zip(paired_feature, feature_len)]

    def forward(
        self,
        paired_wavs: List[Tensor],
        paired_features: Dict[str, Union[Tensor, List[Tensor], Dict[str, Tensor]]],
    ) -> List[Tensor]:
        feature = self._select_feature(paired_features)
        if isinstance(feature, (list, tuple)):
            feature = self._weighted_sum(list(feature))
        return self.tolist(paired_wavs, feature)
