"""GPU tests of the host-side API mirrors: Featurizer (forward + backward of the layer weights) and S3PRLUpstream."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
pytestmark = pytest.mark.gpu


def test_featurizer_forward_backward(s3b_lib):
    import upstream_oracle as O
    from s3prl_b200.hub import hubert_base
    from s3prl_b200.upstream.featurizer import Featurizer

    up = hubert_base().to("cuda")
    feat = Featurizer(up, "hidden_states", upstream_device="cuda").to("cuda")  # as Runner does (runner.py:166-180)
    assert feat.layer_num == 13 and feat.output_dim == 768 and feat.downsample_rate == 320
    g = torch.Generator().manual_seed(3)
    wavs = [torch.randn(n, generator=g).cuda() for n in (16000, 12000, 6400)]
    with torch.no_grad():
        res = up(wavs)
    with torch.no_grad():
        feat.weights.copy_(torch.randn(13, generator=g).cuda())
    out = feat(wavs, res)
    T = res["hidden_states"][0].shape[1]  # 49: slicing f[:round(len/320)] clamps at the frame count, as in the reference
    assert [o.shape[0] for o in out] == [min(n, T) for n in O.featurizer_lengths([16000, 12000, 6400])]
    ref_full = O.weighted_sum([h.cpu() for h in res["hidden_states"]], feat.weights.detach().cpu())
    for o, n, r in zip(out, [49, 38, 20], ref_full):
        assert torch.allclose(o.detach().cpu(), r[:n], atol=1e-5, rtol=1e-5)
    # gradient of the 13 layer weights vs autograd through the plain torch formulation
    loss = sum((o * o).sum() for o in out)
    loss.backward()
    w = feat.weights.detach().clone().requires_grad_(True)
    stacked = torch.stack([h.detach() for h in res["hidden_states"]], 0)
    ws = (torch.softmax(w, -1).view(-1, 1, 1, 1) * stacked).sum(0)
    ref_loss = sum((ws[i, :n] ** 2).sum() for i, n in enumerate([49, 38, 20]))
    ref_loss.backward()
    assert torch.allclose(feat.weights.grad, w.grad, rtol=2e-3, atol=1e-3 * w.grad.abs().max().item())


def test_s3prl_upstream_wrapper(s3b_lib):
    import upstream_oracle as O
    from s3prl_b200.nn import S3PRLUpstream

    model = S3PRLUpstream("hubert_base").to("cuda")
    assert model.num_layers == 13 and model.hidden_sizes == [768] * 13 and model.downsample_rates == [320] * 13
    lens = torch.tensor([16000, 9000, 3200])
    wavs = torch.zeros(3, 16000)
    g = torch.Generator().manual_seed(4)
    for i, n in enumerate(lens.tolist()):
        wavs[i, :n] = torch.randn(n, generator=g)
    all_hs, all_lens = model(wavs.cuda(), lens.cuda())
    assert len(all_hs) == 13
    assert all_lens[0].tolist() == O.s3prl_upstream_lengths(lens.tolist())
    assert all_hs[0].shape == (3, 50, 768)  # len(range(0, 16000, 320)) == 50: last frame repeated once
    assert torch.equal(all_hs[3][:, 49], all_hs[3][:, 48])


def test_nn_featurizer_and_model_wrapper(s3b_lib):
    """s3prl.nn.Featurizer / UpstreamDownstreamModel mirrors (s3prl/nn/upstream.py:234-384) on the device: DistilHuBERT's
    six outputs through S3PRLUpstream, layer selection + normalize through the fused weighted sum, weight gradient vs
    the torch formulation. (The host logic is compared with the reference's own classes in tests/test_host_cpu.py.)"""
    import torch.nn.functional as F

    from s3prl_b200.nn import Featurizer, S3PRLUpstream, UpstreamDownstreamModel

    model = S3PRLUpstream("distilhubert_base").to("cuda")
    assert model.num_layers == 6
    lens = torch.tensor([16000, 9000, 3200])
    wavs = torch.zeros(3, 16000)
    g = torch.Generator().manual_seed(9)
    for i, n in enumerate(lens.tolist()):
        wavs[i, :n] = torch.randn(n, generator=g)
    all_hs, all_lens = model(wavs.cuda(), lens.cuda())
    assert len(all_hs) == 6 and all_hs[0].shape == (3, 50, 768) and all_lens[0].tolist() == [50, 29, 10]
    for sel, norm in ((None, False), ([5, 0, 2], True)):
        feat = Featurizer(model, sel, norm).to("cuda")
        with torch.no_grad():
            feat.weights.copy_(torch.randn(len(feat.weights), generator=g).cuda())
        hs, hs_len = feat(all_hs, all_lens)
        picked = [all_hs[i] for i in feat.layer_selections]
        if norm:
            picked = [F.layer_norm(h, (768,)) for h in picked]
        w = feat.weights.detach().clone().requires_grad_(True)
        ref = (torch.softmax(w, -1).view(-1, 1, 1, 1) * torch.stack(picked, 0)).sum(0)
        assert hs.shape == (3, 50, 768) and torch.equal(hs_len, all_lens[0])
        assert torch.allclose(hs, ref, atol=1e-5, rtol=1e-5)
        hs.square().sum().backward()
        ref.square().sum().backward()
        assert torch.allclose(feat.weights.grad, w.grad, rtol=2e-3, atol=1e-3 * w.grad.abs().max().item())

    class Head(torch.nn.Module):
        output_size = 8

        def __init__(self):
            super().__init__()
            self.proj = torch.nn.Linear(768, 8)

        def forward(self, h, h_len):
            return self.proj(h), h_len

    full = UpstreamDownstreamModel(model, Featurizer(model).to("cuda"), Head().cuda())
    out, out_len = full(wavs.cuda(), lens.cuda())
    assert out.shape == (3, 50, 8) and out_len.tolist() == [50, 29, 10] and out.requires_grad
    assert full.downsample_rate == 320 and full.output_size == 8


def test_frozen_upstream_ctc_training_steps(s3b_lib):
    """The SUPERB recipe of BASELINE config 5 in miniature (s3prl/downstream/runner.py:293-330, ctc/expert.py:64-108):
    frozen upstream under no_grad, trainable Featurizer weights + a linear CTC head, synthetic LibriSpeech-shaped
    batch. The loss must fall and the Featurizer's layer weights must move (gradient through s3b_weighted_sum_backward)."""
    import torch.nn.functional as F
    from torch.nn.utils.rnn import pad_sequence

    from s3prl_b200.hub import hubert_base
    from s3prl_b200.upstream.featurizer import Featurizer

    torch.manual_seed(0)
    up = hubert_base().to("cuda")
    feat = Featurizer(up, "hidden_states", upstream_device="cuda").to("cuda")
    head = torch.nn.Linear(feat.output_dim, 32).cuda()
    opt = torch.optim.Adam(list(feat.parameters()) + list(head.parameters()), lr=3e-3)
    g = torch.Generator().manual_seed(11)
    wavs = [torch.randn(n, generator=g).cuda() for n in (24000, 17000, 9000)]
    labels = [torch.randint(1, 32, (6,), generator=g) for _ in wavs]
    w0 = feat.weights.detach().clone()
    losses = []
    for _ in range(8):
        with torch.no_grad():  # runner.py:300-304: upstream frozen
            res = up(wavs)
        feats = feat(wavs, res)  # list of [T_i, D]
        lens = torch.tensor([f.shape[0] for f in feats])
        logp = F.log_softmax(head(pad_sequence(feats, batch_first=True)), dim=-1).transpose(0, 1)
        loss = F.ctc_loss(logp, pad_sequence(labels, batch_first=True), lens, torch.tensor([6] * len(wavs)), blank=0,
                          zero_infinity=True)
        opt.zero_grad()
        loss.backward()
        assert feat.weights.grad is not None and torch.isfinite(feat.weights.grad).all()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.8 * losses[0], losses
    assert (feat.weights.detach() - w0).abs().max().item() > 1e-4
