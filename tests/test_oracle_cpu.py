"""CPU tests pinning the oracle (oracle/upstream_oracle.py) to golden vectors produced by executing the
reference itself (oracle/make_golden.py -> tests/golden/*.pt). These run without a GPU."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
GOLDEN = ROOT / "tests" / "golden"

import upstream_oracle as O  # noqa: E402
from s3prl_b200.upstream.configs import ARCHS  # noqa: E402
from s3prl_b200.upstream.weights import fabricate_state_dict  # noqa: E402

MODEL_FIXTURES = sorted(p.stem for p in GOLDEN.glob("*.pt") if p.stem in ARCHS)
# full-size 24-layer models take a while on CPU; the fast subset keeps the default CPU suite to a few minutes
FAST = {"hubert_base", "wavlm_base_plus", "wav2vec2_base_960", "unispeech_sat_base_plus", "data2vec_base_960"}


def _wavs(lens, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(n, generator=g) for n in lens]


def test_integer_rules_match_reference():
    fx = torch.load(GOLDEN / "integer_rules.pt", weights_only=False)
    for b in fx["batches"] + fx.get("short_batches", []):
        lens, T = b["lens"], b["T"]
        assert O.conv_output_length(max(lens)) == T
        assert O.valid_frames("hubert", lens, max(lens)) == b["hubert_valid"]
        assert O.valid_frames("wavlm", lens, max(lens)) == b["hubert_valid"]
        assert O.valid_frames("wav2vec2", lens, max(lens)) == b["wav2vec2_valid"]
        assert b["hubert_prefix"]
        assert O.featurizer_lengths(lens) == b["featurizer_len"]
        assert O.s3prl_upstream_lengths(lens) == b["s3prl_upstream_len"]
    got = O.wavlm_relative_bucket(fx["wavlm_rel"], 320, 800)
    assert torch.equal(got, fx["wavlm_bucket"])


def test_native_integer_rules_match_reference(s3b_lib):
    """The C++ bookkeeping of the product (s3b_num_frames / s3b_valid_frames) is integer-exact vs the reference."""
    from s3prl_b200.upstream.expert import UpstreamExpert

    fx = torch.load(GOLDEN / "integer_rules.pt", weights_only=False)
    experts = {
        "hubert": UpstreamExpert(name="hubert_base", state_dict={}),
        "wav2vec2": UpstreamExpert(name="wav2vec2_base_960", state_dict={}),
        "wavlm": UpstreamExpert(name="wavlm_base_plus", state_dict={}),
        "distiller": UpstreamExpert(name="distilhubert_base", state_dict={}),
    }
    for b in fx["batches"] + fx.get("short_batches", []):
        lens, T = b["lens"], b["T"]
        assert experts["hubert"].num_frames(max(lens)) == T
        assert s3b_lib.s3b_num_frames(None, max(lens)) == T
        assert experts["hubert"].valid_frames(lens) == b["hubert_valid"]
        assert experts["wavlm"].valid_frames(lens) == b["hubert_valid"]
        assert experts["wav2vec2"].valid_frames(lens) == b["wav2vec2_valid"], lens
        if min(lens) >= 400:  # the Distiller rule (cal_pad_mask, distiller/model.py:272-286) is pinned through the
            # oracle, which the executed-reference golden of distilhubert_base pins (ragged case in the fixture)
            assert experts["distiller"].valid_frames(lens) == O.valid_frames("distiller", lens, max(lens))
    assert len(fx.get("short_batches", [])) >= 10  # utterances shorter than the receptive field (mask index wraps)


def test_native_wavlm_buckets_match_reference(s3b_lib):
    """The product's own bucket rule (csrc/wavlm.cu wavlm_rel_bucket, host logf) is bit-exact with the table the
    reference's _relative_positions_bucket produced for every relative position in [-2100, 2100] — beyond
    |rel| >= max_distance the bucket saturates, so this is exhaustive for the 320 / 800 configuration."""
    import ctypes as C

    from s3prl_b200 import lib

    fx = torch.load(GOLDEN / "integer_rules.pt", weights_only=False)
    rel = fx["wavlm_rel"].to(torch.int32).contiguous()
    out = torch.empty_like(rel)
    lib.check(s3b_lib.s3b_wavlm_buckets(320, 800, C.cast(rel.data_ptr(), C.POINTER(C.c_int32)), rel.numel(),
                                        C.cast(out.data_ptr(), C.POINTER(C.c_int32))))
    assert torch.equal(out.long(), fx["wavlm_bucket"])
    far = torch.tensor([-100000, -801, -800, 800, 801, 100000], dtype=torch.int32)
    o2 = torch.empty_like(far)
    lib.check(s3b_lib.s3b_wavlm_buckets(320, 800, C.cast(far.data_ptr(), C.POINTER(C.c_int32)), far.numel(),
                                        C.cast(o2.data_ptr(), C.POINTER(C.c_int32))))
    assert o2.tolist() == [159, 159, 159, 319, 319, 319]
    assert torch.equal(o2.long(), O.wavlm_relative_bucket(far.long(), 320, 800))
    assert s3b_lib.s3b_wavlm_buckets(322, 800, None, 0, None) != 0  # rejected, with a message
    assert b"null" in s3b_lib.s3b_last_error() or b"multiple" in s3b_lib.s3b_last_error()


@pytest.mark.parametrize("name", [n for n in MODEL_FIXTURES if n in FAST])
def test_oracle_matches_reference_fast(name):
    _check_model(name)


@pytest.mark.slow
@pytest.mark.parametrize("name", [n for n in MODEL_FIXTURES if n not in FAST])
def test_oracle_matches_reference_large(name):
    _check_model(name)


FULL_SIZE = sorted(p.stem for p in GOLDEN.glob("c[0-9]_*.pt"))


@pytest.mark.slow
@pytest.mark.parametrize("name", FULL_SIZE)
def test_oracle_matches_reference_at_baseline_size(name):
    """BASELINE.json C3 / C4 sequence lengths (T = 999 / 499) on two utterances: the executed reference pins the oracle
    at the sizes the bench runs, not only on the short fixtures."""
    _check_model(name)


def _check_model(name):
    fx = torch.load(GOLDEN / f"{name}.pt", weights_only=False)
    cfg = ARCHS[fx["arch"]]
    sd = fabricate_state_dict(cfg, seed=fx["weight_seed"])
    cs, ts = fx["channel_stride"], fx.get("time_stride", 1)
    for case in fx["cases"]:
        wavs = _wavs(case["lens"], case["wav_seed"])
        with torch.no_grad():
            hs, _pad = O.upstream_forward(wavs, sd, cfg)
        assert len(hs) == case["num_hidden"]
        assert tuple(hs[0].shape) == tuple(case["shape"])
        for l, h in enumerate(hs):
            ref = case["sub"][l]
            got = h[:, (h.shape[1] - 1) % ts :: ts, ::cs]
            rel = ((got.double() - ref.double()).norm() / ref.double().norm()).item()
            # two fp32 CPU evaluations of the same math (different op order / BLAS blocking)
            assert rel < 2e-5, (name, l, rel)
            assert abs(h.double().norm().item() - case["norms"][l].item()) < 2e-5 * case["norms"][l].item()


def test_gelu_formula():
    """fp32 emulation of the device GELU (s3prl_b200/csrc/common.cuh gelu_erf: erfc by Abramowitz-Stegun 7.1.26)
    against the float64 exact-erf GELU the reference computes in fp32 (nn.GELU(), wav2vec2_model.py:2893)."""
    import math

    import numpy as np

    f = np.float32
    x = np.concatenate([np.linspace(-12, 12, 400001), np.random.default_rng(0).standard_normal(200000) * 2]).astype(f)
    z = (np.abs(x) * f(0.70710678118654752440)).astype(f)
    t = (f(1) / (f(0.3275911) * z + f(1)).astype(f)).astype(f)
    poly = f(1.061405429) * t + f(-1.453152027)
    for c in (1.421413741, -0.284496736, 0.254829592):
        poly = (poly * t + f(c)).astype(f)
    e = np.exp2((z * (z * f(-1.4426950408889634))).astype(f)).astype(f)
    erfc_z = (poly * t * e).astype(f)
    out = (f(0.5) * x * np.where(x >= 0, f(2) - erfc_z, erfc_z)).astype(f)
    xd = x.astype(np.float64)
    ref = 0.5 * xd * (1 + np.vectorize(math.erf)(xd / math.sqrt(2)))
    assert np.abs(out - ref).max() < 1e-6
    torch_fp32 = torch.nn.functional.gelu(torch.from_numpy(x)).numpy()
    assert np.abs(out - ref).max() <= np.abs(torch_fp32 - ref).max()  # at least as close to exact as torch's fp32 GELU


def test_native_integer_rules_property(s3b_lib):
    """Randomised (hypothesis) agreement of the C++ frame bookkeeping with the oracle's mask construction (which is
    pinned to the reference by integer_rules.pt): ragged batches, boundary lengths around multiples of 320 and of the
    400-sample receptive field, single-utterance batches, all three mask rules."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from s3prl_b200.upstream.expert import UpstreamExpert

    experts = {
        "hubert": UpstreamExpert(name="hubert_base", state_dict={}),
        "wav2vec2": UpstreamExpert(name="wav2vec2_base_960", state_dict={}),
        "wavlm": UpstreamExpert(name="wavlm_base_plus", state_dict={}),
        "distiller": UpstreamExpert(name="distilhubert_base", state_dict={}),
    }
    length = st.one_of(
        st.integers(min_value=400, max_value=200000),
        st.builds(lambda k, d: max(400, 320 * k + d), st.integers(1, 600), st.integers(-3, 83)),
    )

    @settings(max_examples=150, deadline=None)
    @given(st.lists(length, min_size=1, max_size=6))
    def check(lens):
        T = O.conv_output_length(max(lens))
        assert s3b_lib.s3b_num_frames(None, max(lens)) == T
        for fam, ex in experts.items():
            assert ex.valid_frames(lens) == O.valid_frames(fam, lens, max(lens)), (fam, lens)

    check()
