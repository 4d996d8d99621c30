"""TEST INFRASTRUCTURE — CPU oracle of the upstream feature-extraction hot path.

A plain restatement, in fp32 torch-CPU functional ops and integer numpy/python arithmetic, of what the
reference computes for ``UpstreamExpert(wavs)["hidden_states"]``. Every function cites
the reference lines it follows. Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module; the product package never does.

Pinning: ``tests/test_oracle_cpu.py`` checks this file against golden vectors produced by *executing the
reference itself* (``oracle/make_golden.py`` imports /root/reference and saves its outputs under
``tests/golden``). The reference publishes no offline golden vectors for this path (they are network-only,
test/test_upstream.py:25-26), so for pretrained weights parity is unpinned; for the fabricated checkpoints it
is pinned by those executed-reference fixtures.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

CONV_LAYERS = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2


# ------------------------------------------------------------------------------------------------
# integer bookkeeping (bit-exact rules)
# ------------------------------------------------------------------------------------------------
def conv_output_length(n: int) -> int:
    """ConvFeatureExtractionModel output length, L_out = floor((L - k) / s) + 1 per layer
    (s3prl/upstream/wav2vec2/wav2vec2_model.py:2857-2934)."""
    for _d, k, s in CONV_LAYERS:
        n = (n - k) // s + 1 if n >= k else 0
    return n


def sample_padding_mask(lens: Sequence[int], max_len: int) -> torch.Tensor:
    """pad[b, n] = n >= len_b (s3prl/upstream/hubert/expert.py:61-65)."""
    return ~torch.lt(torch.arange(max_len).unsqueeze(0), torch.tensor(list(lens), dtype=torch.long).unsqueeze(1))


def frame_padding_mask_chunk_all(sample_mask: torch.Tensor, T: int) -> torch.Tensor:
    """HuBERT / WavLM ``forward_padding_mask`` (s3prl/upstream/hubert/hubert_model.py:454-464,
    s3prl/upstream/wavlm/WavLM.py:339-349): drop Lmax % T tail samples, view [B, T, Lmax // T], all(-1)."""
    extra = sample_mask.size(1) % T
    if extra > 0:
        sample_mask = sample_mask[:, :-extra]
    return sample_mask.view(sample_mask.size(0), T, -1).all(-1)


def frame_padding_mask_conv_length(sample_mask: torch.Tensor, T: int) -> Optional[torch.Tensor]:
    """wav2vec 2.0 rule (s3prl/upstream/wav2vec2/wav2vec2_model.py:2610-2625, 2652-2671): ``None`` when the batch
    has no padded sample; otherwise valid length through the conv formula in *float* arithmetic followed by
    the scatter + flip/cumsum/flip trick."""
    if not bool(sample_mask.any()):
        return None
    input_lengths = (1 - sample_mask.long()).sum(-1)
    out = input_lengths
    for _d, k, s in CONV_LAYERS:
        out = torch.floor((out - k) / s + 1)
    out = out.to(torch.long)
    mask = torch.zeros((sample_mask.size(0), T), dtype=torch.float32)
    mask[(torch.arange(mask.shape[0]), out - 1)] = 1
    return (1 - mask.flip([-1]).cumsum(-1).flip([-1])).bool()


def valid_frames(family: str, lens: Sequence[int], max_len: int) -> List[int]:
    """Un-padded frame count per utterance implied by the family's frame mask rule."""
    T = conv_output_length(max_len)
    if family == "distiller":
        return distiller_valid_frames(lens, T)
    sm = sample_padding_mask(lens, max_len)
    conv_rule = family in ("wav2vec2", "data2vec")  # data2vec_model.py:455-476 repeats the wav2vec 2.0 rule
    fm = frame_padding_mask_conv_length(sm, T) if conv_rule else frame_padding_mask_chunk_all(sm, T)
    if fm is None:
        return [T] * len(lens)
    return [int((~row).sum()) for row in fm]


def featurizer_lengths(lens: Sequence[int], rate: int = 320) -> List[int]:
    """Featurizer.tolist: round(len / downsample_rate), Python banker's rounding (interfaces.py:250-261)."""
    return [round(n / rate) for n in lens]


def s3prl_upstream_lengths(lens: Sequence[int], rate: int = 320) -> List[int]:
    """S3PRLUpstream: h_len = (len - 1) // stride + 1 (s3prl/nn/upstream.py:223)."""
    return [(n - 1) // rate + 1 for n in lens]


def wavlm_relative_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """_relative_positions_bucket, bidirectional (s3prl/upstream/wavlm/modules.py:418-448)."""
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).to(
        torch.long
    )
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


# ------------------------------------------------------------------------------------------------
# floating-point path (fp32)
# ------------------------------------------------------------------------------------------------
def pad_waveforms(wavs: Sequence[torch.Tensor], normalize: bool, max_len: Optional[int] = None) -> torch.Tensor:
    """Optional F.layer_norm(wav, wav.shape) per utterance then zero padding
    (s3prl/upstream/hubert/expert.py:57-66)."""
    if normalize:
        wavs = [F.layer_norm(w, w.shape) for w in wavs]
    L = max_len or max(len(w) for w in wavs)
    out = torch.zeros(len(wavs), L, dtype=torch.float32)
    for i, w in enumerate(wavs):
        out[i, : len(w)] = w
    return out


def conv_feature_extractor(x: torch.Tensor, sd: Dict[str, torch.Tensor], mode: str, conv_bias: bool) -> torch.Tensor:
    """[B, L] -> [B, 512, T] (s3prl/upstream/wav2vec2/wav2vec2_model.py:2869-2934; GroupNorm/LayerNorm in fp32
    :1830-1853). GroupNorm statistics run over the whole padded length."""
    x = x.unsqueeze(1)
    for i, (_dim, _k, s) in enumerate(CONV_LAYERS):
        p = f"feature_extractor.conv_layers.{i}"
        x = F.conv1d(x, sd[f"{p}.0.weight"], sd.get(f"{p}.0.bias") if conv_bias else None, stride=s)
        if mode == "layer_norm":
            x = F.layer_norm(x.transpose(-2, -1), (x.size(1),), sd[f"{p}.2.1.weight"], sd[f"{p}.2.1.bias"], 1e-5)
            x = x.transpose(-2, -1)
        elif i == 0:
            x = F.group_norm(x, x.size(1), sd[f"{p}.2.weight"], sd[f"{p}.2.bias"], 1e-5)
        x = F.gelu(x)
    return x


def positional_conv(x: torch.Tensor, sd: Dict[str, torch.Tensor], groups: int) -> torch.Tensor:
    """x: [B, T, D] -> GELU(SamePad(Conv1d(weight_norm(dim=2)))) (wav2vec2_model.py:2937-2953, 1803-1808)."""
    v, g = sd["encoder.pos_conv.0.weight_v"], sd["encoder.pos_conv.0.weight_g"]
    k = v.size(2)
    w = v * (g / v.norm(dim=(0, 1), keepdim=True))
    y = F.conv1d(x.transpose(1, 2), w, sd["encoder.pos_conv.0.bias"], padding=k // 2, groups=groups)
    if k % 2 == 0:
        y = y[:, :, :-1]
    return F.gelu(y).transpose(1, 2)


def positional_conv_blocks(x: torch.Tensor, sd: Dict[str, torch.Tensor], groups: int, depth: int) -> torch.Tensor:
    """data2vec (pos_conv_depth > 1): x: [B, T, D] -> depth x [Conv1d(k, padding k // 2, groups) -> SamePad(k) ->
    LayerNorm over the channels without affine -> GELU] (make_conv_block, wav2vec2_model.py:2995-3026). Nothing is
    re-masked between the blocks."""
    y = x.transpose(1, 2)
    for i in range(depth):
        w = sd[f"encoder.pos_conv.{i}.0.weight"]
        k = w.size(2)
        y = F.conv1d(y, w, sd[f"encoder.pos_conv.{i}.0.bias"], padding=k // 2, groups=groups)
        if k % 2 == 0:
            y = y[:, :, :-1]
        y = F.layer_norm(y.transpose(1, 2), (y.size(1),), None, None, 1e-5).transpose(1, 2)
        y = F.gelu(y)
    return y.transpose(1, 2)


def self_attention(
    x: torch.Tensor,
    sd: Dict[str, torch.Tensor],
    prefix: str,
    heads: int,
    key_padding_mask: Optional[torch.Tensor],
    attn_bias: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """F.multi_head_attention_forward semantics (wav2vec2_model.py:1146-1168): separate q/k/v weights, q scaled
    by head_dim**-0.5, masked keys get -inf, softmax over keys in fp32, padded QUERY rows are still computed.
    attn_bias: optional float [B, H, T, T] added to the scaled scores (WavLM)."""
    B, T, D = x.shape
    hd = D // heads
    q = F.linear(x, sd[f"{prefix}.q_proj.weight"], sd[f"{prefix}.q_proj.bias"])
    k = F.linear(x, sd[f"{prefix}.k_proj.weight"], sd[f"{prefix}.k_proj.bias"])
    v = F.linear(x, sd[f"{prefix}.v_proj.weight"], sd[f"{prefix}.v_proj.bias"])
    q = q.view(B, T, heads, hd).transpose(1, 2) * (hd**-0.5)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if attn_bias is not None:
        s = s + attn_bias
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, T, D)
    return F.linear(o, sd[f"{prefix}.out_proj.weight"], sd[f"{prefix}.out_proj.bias"])


def wavlm_position_bias(sd: Dict[str, torch.Tensor], T: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """compute_bias (s3prl/upstream/wavlm/modules.py:450-462): [H, T, T] from layer 0's embedding table."""
    ctx = torch.arange(T, dtype=torch.long)[:, None]
    mem = torch.arange(T, dtype=torch.long)[None, :]
    bucket = wavlm_relative_bucket(mem - ctx, num_buckets, max_distance)
    emb = sd["encoder.layers.0.self_attn.relative_attention_bias.weight"]
    return F.embedding(bucket, emb).permute(2, 0, 1)


def wavlm_gated_bias(x: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, heads: int, bias: torch.Tensor) -> torch.Tensor:
    """gru_rel_pos gate (s3prl/upstream/wavlm/modules.py:534-551): x [B, T, D] (the attention input)."""
    B, T, D = x.shape
    xh = x.view(B, T, heads, -1).permute(0, 2, 1, 3)  # [B, H, T, 64]
    u = F.linear(xh, sd[f"{prefix}.grep_linear.weight"], sd[f"{prefix}.grep_linear.bias"])
    gate_a, gate_b = torch.sigmoid(u.view(B, heads, T, 2, 4).sum(-1)).chunk(2, dim=-1)
    gate_a_1 = gate_a * (gate_b * sd[f"{prefix}.grep_a"] - 1.0) + 2.0  # [B, H, T, 1]
    return gate_a_1 * bias.unsqueeze(0)


def upstream_forward(
    wavs: Sequence[torch.Tensor],
    sd: Dict[str, torch.Tensor],
    cfg,
    max_len: Optional[int] = None,
) -> Tuple[List[torch.Tensor], Optional[torch.Tensor]]:
    """hidden_states (list of NL+1 tensors [B, T, D]) and the frame padding mask, for an ``ArchConfig``-like
    ``cfg`` (fields: family, extractor_mode, conv_bias, layer_norm_first, normalize, encoder_layers,
    encoder_attention_heads, conv_pos_groups, relative_position_embedding, num_buckets, max_distance, gru_rel_pos).

    Follows HubertModel.forward (hubert_model.py:466-513), Wav2Vec2Model.forward (wav2vec2_model.py:2638-2684),
    Data2VecAudioModel.forward with features_only (data2vec_model.py:428-560: the wav2vec 2.0 steps, no masking),
    WavLM.extract_features (WavLM.py:351-405), TransformerEncoder.extract_features (wav2vec2_model.py:3054-3121;
    WavLM.py:599-645) and the layer forward (wav2vec2_model.py:3260-3322; WavLM.py:709-774). The reference's
    pad-to-multiple-of-2 column (wav2vec2_model.py:3073-3082) is a masked key whose query row is discarded by
    hook_postprocess (hubert/expert.py:45-51): it cannot influence the returned frames and is omitted.
    """
    if getattr(cfg, "family", None) == "distiller":
        return distiller_forward(wavs, sd, cfg, max_len)
    sd = {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}
    wavs = [w.float() for w in wavs]
    lens = [len(w) for w in wavs]
    x = pad_waveforms(wavs, cfg.normalize, max_len)
    L = x.size(1)
    sample_mask = sample_padding_mask(lens, L)
    feats = conv_feature_extractor(x, sd, cfg.extractor_mode, cfg.conv_bias)  # [B, 512, T]
    feats = feats.transpose(1, 2)
    T = feats.size(1)
    feats = F.layer_norm(feats, (feats.size(-1),), sd["layer_norm.weight"], sd["layer_norm.bias"], 1e-5)
    if cfg.family in ("wav2vec2", "data2vec"):
        pad = frame_padding_mask_conv_length(sample_mask, T)
    else:
        pad = frame_padding_mask_chunk_all(sample_mask, T)
    h = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    if pad is not None:
        h = h.masked_fill(pad.unsqueeze(-1), 0.0)
    depth = int(getattr(cfg, "pos_conv_depth", 1))
    if depth > 1:
        h = h + positional_conv_blocks(h, sd, cfg.conv_pos_groups, depth)
    else:
        h = h + positional_conv(h, sd, cfg.conv_pos_groups)
    if not cfg.layer_norm_first:
        h = F.layer_norm(h, (h.size(-1),), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5)

    heads = cfg.encoder_attention_heads
    rel = bool(getattr(cfg, "relative_position_embedding", False))
    pos_bias = wavlm_position_bias(sd, T, cfg.num_buckets, cfg.max_distance) if rel else None

    def ln(t, name):
        return F.layer_norm(t, (t.size(-1),), sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-5)

    hidden = []
    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}"
        hidden.append(h)
        attn_in = ln(h, f"{p}.self_attn_layer_norm") if cfg.layer_norm_first else h
        bias = None
        if rel:
            bias = wavlm_gated_bias(attn_in, sd, f"{p}.self_attn", heads, pos_bias) if cfg.gru_rel_pos else pos_bias.unsqueeze(0)
        a = self_attention(attn_in, sd, f"{p}.self_attn", heads, pad, bias)
        if cfg.layer_norm_first:
            h = h + a
            f = ln(h, f"{p}.final_layer_norm")
            f = F.linear(F.gelu(F.linear(f, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"])), sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"])
            h = h + f
        else:
            h = ln(h + a, f"{p}.self_attn_layer_norm")
            f = F.linear(F.gelu(F.linear(h, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"])), sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"])
            h = ln(h + f, f"{p}.final_layer_norm")
    if cfg.layer_norm_first:
        h = ln(h, "encoder.layer_norm")  # TransformerEncoder.forward, wav2vec2_model.py:3049-3050
    hidden.append(h)
    return hidden, pad


def distiller_valid_frames(lens: Sequence[int], T: int) -> List[int]:
    """DistillerModel.cal_pad_mask (s3prl/upstream/distiller/model.py:272-286): the conv length formula with
    truncating division on every utterance (padded batch or not); frames >= that length are padding."""
    out = []
    for n in lens:
        for _d, k, s_ in CONV_LAYERS:
            n = int((n - k) / s_) + 1  # torch.div(..., rounding_mode="trunc")
        out.append(max(0, min(n, T)) if n >= 0 else max(0, T + n))  # new_pad_mask[idx, n:] = 0 (Python slice semantics)
    return out


def distiller_forward(wavs: Sequence[torch.Tensor], sd: Dict[str, torch.Tensor], cfg, max_len: Optional[int] = None):
    """hidden_states of the Distiller expert (s3prl/upstream/distiller/expert.py:44-63):
    [feat_final] + layer outputs + prediction heads, and the frame padding mask.
    Follows DistillerModel.forward (distiller/model.py:187-269: no LayerNorm in front of post_extract_proj; the
    encoder zeroes the padded frames of feat_final IN PLACE, so the returned feat_final carries the zeros), the
    TransformerEncoder of distiller/module.py:292-334 (layer outputs are collected) and the output layer
    Linear -> GELU -> SplitLinear (model.py:150-160, module.py:55-90)."""
    sd = {k: v.float() for k, v in sd.items() if torch.is_tensor(v) and v.is_floating_point()}
    wavs = [w.float() for w in wavs]
    lens = [len(w) for w in wavs]
    x = pad_waveforms(wavs, False, max_len)
    feats = conv_feature_extractor(x, sd, cfg.extractor_mode, False).transpose(1, 2)  # [B, T, 512]
    B, T, _ = feats.shape
    valid = distiller_valid_frames(lens, T)
    pad = torch.arange(T).unsqueeze(0) >= torch.tensor(valid).unsqueeze(1)
    feat_final = F.linear(feats, sd["post_extract_proj.weight"], sd["post_extract_proj.bias"])
    feat_final = feat_final.masked_fill(pad.unsqueeze(-1), 0.0)
    h = feat_final + positional_conv(feat_final, sd, cfg.conv_pos_groups)
    if not cfg.layer_norm_first:
        h = F.layer_norm(h, (h.size(-1),), sd["encoder.layer_norm.weight"], sd["encoder.layer_norm.bias"], 1e-5)

    def ln(t, name):
        return F.layer_norm(t, (t.size(-1),), sd[f"{name}.weight"], sd[f"{name}.bias"], 1e-5)

    heads = cfg.encoder_attention_heads
    layer_out = []
    for l in range(cfg.encoder_layers):
        p = f"encoder.layers.{l}"
        if cfg.layer_norm_first:
            a = self_attention(ln(h, f"{p}.self_attn_layer_norm"), sd, f"{p}.self_attn", heads, pad)
            h = h + a
            f = ln(h, f"{p}.final_layer_norm")
            h = h + F.linear(F.gelu(F.linear(f, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"])), sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"])
        else:
            a = self_attention(h, sd, f"{p}.self_attn", heads, pad)
            h = ln(h + a, f"{p}.self_attn_layer_norm")
            f = F.linear(F.gelu(F.linear(h, sd[f"{p}.fc1.weight"], sd[f"{p}.fc1.bias"])), sd[f"{p}.fc2.weight"], sd[f"{p}.fc2.bias"])
            h = ln(h + f, f"{p}.final_layer_norm")
        layer_out.append(h)
    hidden = ln(h, "encoder.layer_norm") if cfg.layer_norm_first else h
    N = cfg.pred_heads
    z = F.gelu(F.linear(hidden, sd["output_layer.0.weight"], sd["output_layer.0.bias"]))  # [B, T, N*D]
    D = hidden.size(-1)
    z = z.reshape(B, T, N, 1, D)
    pred = torch.einsum("...klm,kmn->...kln", z, sd["output_layer.2.weight"]).squeeze(3) + sd["output_layer.2.bias"]
    preds = [pred[:, :, i, :] for i in range(N)]
    return [feat_final] + layer_out + preds, pad


def weighted_sum(hidden: Sequence[torch.Tensor], weights: torch.Tensor) -> torch.Tensor:
    """Featurizer._weighted_sum (s3prl/upstream/interfaces.py:217-248), normalize=False."""
    stacked = torch.stack(list(hidden), dim=0)
    w = F.softmax(weights, dim=-1)
    return (w.view(-1, *([1] * (stacked.dim() - 1))) * stacked).sum(dim=0)
