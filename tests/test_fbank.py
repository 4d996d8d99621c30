"""fbank baseline (BASELINE.json configs[0]): oracle pinned to the executed reference (CPU), CUDA kernel vs both (GPU).
Tolerance: compare after CMVN with atol 1e-3 (fp32 FFT/log path)."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
GOLDEN = ROOT / "tests" / "golden" / "fbank.pt"


def _wavs(case):
    if case["name"].startswith("c1"):
        torch.manual_seed(0)
        return [torch.randn(16000) for _ in range(4)]
    g = torch.Generator().manual_seed(321)
    return [torch.randn(n, generator=g) for n in case["lens"]]


def test_fbank_oracle_matches_reference():
    import fbank_oracle as FO

    fx = torch.load(GOLDEN, weights_only=False)
    for case in fx["cases"]:
        got = FO.fbank_forward(_wavs(case))
        assert got.shape == case["out"].shape
        assert FO.num_frames(16000) == 98
        assert torch.allclose(got, case["out"], atol=2e-4, rtol=0, equal_nan=True)  # 1-frame utterance: std is NaN in the reference too


@pytest.mark.gpu
def test_fbank_cuda_matches_reference_and_oracle(s3b_lib):
    import fbank_oracle as FO
    from s3prl_b200.hub import fbank

    expert = fbank().to("cuda")
    assert expert.get_downsample_rates("hidden_states") == 160
    fx = torch.load(GOLDEN, weights_only=False)
    for case in fx["cases"]:
        wavs = _wavs(case)
        res = expert([w.cuda() for w in wavs])
        got = res["hidden_states"][0].cpu()
        assert res["last_hidden_state"].shape == case["out"].shape
        assert torch.equal(torch.isnan(got), torch.isnan(case["out"]))
        err = (got - case["out"]).nan_to_num().abs().max().item()
        print(f"fbank {case['name']}: max abs err vs reference golden {err:.3e}")
        assert err < 1e-3
    # larger ragged batch vs the oracle
    g = torch.Generator().manual_seed(11)
    wavs = [torch.randn(n, generator=g) for n in (160000, 80000, 31234, 400)]
    ref = FO.fbank_forward(wavs[:3])
    got = expert([w.cuda() for w in wavs[:3]])["hidden_states"][0].cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-3


# ---- mel / linear (torch.stft path) ---------------------------------------------------------------------------
SPEC_GOLDEN = ROOT / "tests" / "golden" / "spectrogram.pt"


def _spec_wavs(case):
    if case["name"].startswith("4x1s"):
        torch.manual_seed(0)
        return [torch.randn(16000) for _ in range(4)]
    g = torch.Generator().manual_seed(654)
    wavs = [torch.randn(n, generator=g) for n in case["lens"]]
    wavs[1][-37:] = 0.0
    return wavs


def test_spectrogram_oracle_matches_reference():
    import fbank_oracle as FO

    fx = torch.load(SPEC_GOLDEN, weights_only=False)
    for case in fx["cases"]:
        got = FO.spectrogram_forward(_spec_wavs(case), case["feat"])
        assert got.shape == case["out"].shape, (case["feat"], case["name"], got.shape)
        assert torch.allclose(got, case["out"], atol=3e-4, rtol=0), (got - case["out"]).abs().max()


def test_spectrogram_length_plan_matches_reference():
    from s3prl_b200.upstream.baseline import SpectrogramExpert

    fx = torch.load(SPEC_GOLDEN, weights_only=False)
    for case in fx["cases"]:
        wavs = _spec_wavs(case)
        trimmed = [len(w) if len(w.nonzero()) == 0 else int(w.nonzero()[:, -1].max()) + 1 for w in wavs]
        lp, feats_len, final_len, t_out = SpectrogramExpert.plan_lengths([len(w) for w in wavs], trimmed)
        assert t_out == case["out"].shape[1]
        for b, n in enumerate(final_len):  # frames past the kept length are exact zeros in the reference output
            assert torch.count_nonzero(case["out"][b, n:]) == 0
            assert n == 0 or torch.count_nonzero(case["out"][b, n - 1]) > 0


@pytest.mark.gpu
def test_spectrogram_cuda_matches_reference(s3b_lib):
    from s3prl_b200 import hub

    fx = torch.load(SPEC_GOLDEN, weights_only=False)
    experts = {"mel": hub.mel().to("cuda"), "linear": hub.linear().to("cuda")}
    for case in fx["cases"]:
        wavs = _spec_wavs(case)
        res = experts[case["feat"]]([w.cuda() for w in wavs])
        got = res["hidden_states"][0].cpu()
        assert got.shape == case["out"].shape
        err = (got - case["out"]).abs().max().item()
        print(f"{case['feat']} {case['name']}: max abs err vs reference golden {err:.3e}")
        assert err < 1e-3
