"""TEST INFRASTRUCTURE — CPU emulation of candidate tensor-core operand schemes.

Answers "how much of the 1e-3 budget would a cheaper MMA scheme spend?" without GPU time: the oracle forward is run
with every contraction weight / activation rounded the way the candidate scheme would round it, and each hidden state
is compared with the exact fp32 oracle (relative Frobenius error per layer).

  bf16x3   : A = A_hi + A_lo, W = W_hi + W_lo (bf16), products hi*hi + hi*lo + lo*hi  (the shipped scheme; 3 MMAs)
  fp16x2   : A = A_hi + A_lo (fp16), W rounded ONCE to fp16, products A_hi*W_hi + A_lo*W_hi           (2 MMAs)
  bf16x3f8 : bf16 hi*hi + the two correction products with e4m3 operands (4-bit significands)       (2 MMA units)
  fp16x2f8 : fp16x2 + the dropped A_hi*W_lo product with e4m3 operands                               (2.5 MMA units)
  fp16f8x2 : fp16 hi*hi + BOTH first-order corrections (A_lo*W, A*W_lo) with e4m3 operands               (2 MMA units)
  tf32     : both operands rounded to 11-bit significands                                             (2 MMA units)

    python oracle/numerics_emulation.py [--models hubert_base wav2vec2_large_ll60k ...] [--seconds 2]
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import upstream_oracle as O  # noqa: E402
from s3prl_b200.upstream.configs import ARCHS  # noqa: E402
from s3prl_b200.upstream.weights import fabricate_state_dict  # noqa: E402


def _round_sig(x: torch.Tensor, bits: int) -> torch.Tensor:
    """round-to-nearest-even to a `bits`-bit significand (no exponent-range limit)."""
    m, e = torch.frexp(x.double())
    return torch.ldexp(torch.round(m * (1 << bits)) / (1 << bits), e).float()


def _bf16(x):
    return x.to(torch.bfloat16).float()


def _fp16(x):
    return x.to(torch.float16).float()


def contraction(scheme: str, a: torch.Tensor, w: torch.Tensor, op):
    """op(a, w) is the exact fp32 contraction (linear / conv); returns the value the scheme would produce (products
    summed in fp32/fp64 by torch — the accumulation itself is not the object of this experiment)."""
    if scheme == "exact":
        return op(a, w)
    if scheme == "bf16x3":
        ah, wh = _bf16(a), _bf16(w)
        al, wl = _bf16(a - ah), _bf16(w - wh)
        return op(ah, wh) + op(ah, wl) + op(al, wh)
    if scheme == "fp16x2":
        ah, wh = _fp16(a), _fp16(w)
        al = _fp16(a - ah)
        return op(ah, wh) + op(al, wh)
    if scheme == "bf16x3f8":
        ah, wh = _bf16(a), _bf16(w)
        al, wl = _bf16(a - ah), _bf16(w - wh)
        q = lambda t: _round_sig(t, 4)  # e4m3: 4-bit significand (range handled by power-of-two scaling)
        return op(ah, wh) + op(q(ah), q(wl)) + op(q(al), q(wh))
    if scheme == "fp16x2f8":
        ah, wh = _fp16(a), _fp16(w)
        al, wl = _fp16(a - ah), _fp16(w - wh)
        q = lambda t: _round_sig(t, 4)
        return op(ah, wh) + op(al, wh) + op(q(ah), q(wl))
    if scheme == "fp16f8x2":
        ah, wh = _fp16(a), _fp16(w)
        al, wl = a - ah, w - wh
        q = lambda t: _round_sig(t, 4)
        return op(ah, wh) + op(q(al), q(wh)) + op(q(ah), q(wl))
    if scheme == "tf32":
        return op(_round_sig(a, 11), _round_sig(w, 11))
    raise ValueError(scheme)


class Patched:
    """Route F.linear and the k>1 convolutions of the oracle through `contraction` (conv-0 stays fp32: CUDA cores)."""

    def __init__(self, scheme):
        self.scheme = scheme

    def __enter__(self):
        self.lin, self.conv = F.linear, F.conv1d
        sch = self.scheme

        def linear(x, w, b=None):
            y = contraction(sch, x, w, lambda a, ww: self.lin(a, ww))
            return y if b is None else y + b

        def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if w.shape[1] == 1:  # conv-0
                return self.conv(x, w, b, stride, padding, dilation, groups)
            y = contraction(sch, x, w, lambda a, ww: self.conv(a, ww, None, stride, padding, dilation, groups))
            return y if b is None else y + b.view(1, -1, 1)

        F.linear, F.conv1d = linear, conv1d
        # attention score / PV products: emulate on q,k / p,v through torch.matmul is left exact (both operands are
        # activations; the shipped kernel splits both) — the weight contractions dominate the error budget
        return self

    def __exit__(self, *a):
        F.linear, F.conv1d = self.lin, self.conv


def run(models, seconds, schemes):
    rows = []
    for name in models:
        cfg = ARCHS[name]
        sd = fabricate_state_dict(cfg, 0)
        g = torch.Generator().manual_seed(42)
        wavs = [torch.randn(int(seconds * 16000), generator=g), torch.randn(int(seconds * 16000 * 0.7), generator=g)]
        with torch.no_grad():
            ref, _ = O.upstream_forward(wavs, sd, cfg)
            for sch in schemes:
                with Patched(sch):
                    got, _ = O.upstream_forward(wavs, sd, cfg)
                errs = [((a.double() - b.double()).norm() / b.double().norm()).item() for a, b in zip(got, ref)]
                rows.append((name, sch, max(errs), errs[0], errs[len(errs) // 2], errs[-1]))
                print(f"{name:24s} {sch:9s} worst {max(errs):.3e}  first {errs[0]:.3e}  mid {errs[len(errs)//2]:.3e}  last {errs[-1]:.3e}",
                      flush=True)
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", nargs="+", default=["hubert_base", "wav2vec2_large_ll60k", "wav2vec2_large_960"])
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--schemes", nargs="+", default=["bf16x3", "fp16x2", "bf16x3f8", "tf32"])
    a = ap.parse_args()
    torch.set_num_threads(8)
    run(a.models, a.seconds, a.schemes)
