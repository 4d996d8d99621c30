"""GPU parity tests of the building-block entry points of the C ABI (tcgen05 GEMM, LayerNorm, attention)
against fp64 torch math on the same inputs. Tolerance: the north_star bar is 1e-3 relative on hidden
states; single blocks are held to 1e-4 relative Frobenius / 2e-4 of the output scale elementwise."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize(
    "M,N,K,bias,gelu,res",
    [
        (128, 256, 64, False, False, False),
        (128, 256, 256, True, False, False),
        (300, 768, 512, True, False, True),
        (1000, 2304, 768, True, True, False),
        (4099, 3072, 768, True, True, True),
        (257, 128, 1024, False, False, False),
        (130, 64, 128, True, False, False),
        (64, 48, 192, True, True, False),
        (2000, 768, 3072, True, False, True),
    ],
)
def test_linear_bf16x3(s3b_lib, M, N, K, bias, gelu, res):
    from s3prl_b200 import lib

    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K**0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    r = torch.randn(M, N, device="cuda", generator=g) if res else None
    out = torch.full((M, N), float("nan"), device="cuda")
    lib.check(
        s3b_lib.s3b_linear_f32(
            _ptr(a), _ptr(w), _ptr(b) if bias else None, _ptr(r) if res else None, M, N, K, int(gelu), _ptr(out), _stream()
        )
    )
    torch.cuda.synchronize()
    ref = a.double() @ w.double().t()
    if bias:
        ref = ref + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + r.double()
    assert torch.isfinite(out).all()
    rel = _rel(out, ref)
    mx = (out.double() - ref).abs().max().item()
    print(f"linear M={M} N={N} K={K}: rel={rel:.3e} maxabs={mx:.3e}")
    assert rel < 1e-4, rel
    assert mx < 2e-4 * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("D", [512, 768, 1024])
@pytest.mark.parametrize("gelu", [0, 1])
def test_layernorm(s3b_lib, D, gelu):
    from s3prl_b200 import lib

    M = 1037
    x = torch.randn(M, D, device="cuda") * 3 + 0.5
    g = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    out = torch.empty_like(x)
    lib.check(s3b_lib.s3b_layernorm_f32(_ptr(x), M, D, _ptr(g), _ptr(b), gelu, _ptr(out), _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.double(), (D,), g.double(), b.double(), 1e-5)
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert _rel(out, ref) < 2e-6


@pytest.mark.parametrize(
    "B,T,H,valid",
    [
        (1, 64, 1, [64]),
        (2, 200, 2, [200, 130]),
        (3, 499, 12, [499, 250, 1]),
        (2, 129, 4, [129, 65]),
        # BASELINE C3 length (wav2vec2_large 20 s -> T = 999: 16 key blocks, V^T row stride Tp = 1000) and neighbours
        (2, 999, 16, [999, 640]),
        (2, 1000, 2, [1000, 961]),
        (3, 1023, 2, [1023, 1, 513]),
        (1, 1025, 1, [1025]),
    ],
)
def test_attention(s3b_lib, B, T, H, valid):
    from s3prl_b200 import lib

    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * 100 + T)
    q = torch.randn(B, T, D, device="cuda", generator=g)
    k = torch.randn(B, T, D, device="cuda", generator=g)
    v = torch.randn(B, T, D, device="cuda", generator=g)
    out = torch.full((B, T, D), float("nan"), device="cuda")
    vf = (C.c_int32 * B)(*valid)
    lib.check(s3b_lib.s3b_attention_f32(_ptr(q), _ptr(k), _ptr(v), vf, B, T, H, _ptr(out), _stream()))
    torch.cuda.synchronize()
    qh = q.double().view(B, T, H, 64).transpose(1, 2)
    kh = k.double().view(B, T, H, 64).transpose(1, 2)
    vh = v.double().view(B, T, H, 64).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) * 0.125
    mask = torch.arange(T, device="cuda")[None, :] >= torch.tensor(valid, device="cuda")[:, None]
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    ref = (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, T, D)
    assert torch.isfinite(out).all()
    rel = _rel(out, ref)
    print(f"attention B={B} T={T} H={H}: rel={rel:.3e}")
    assert rel < 1e-4, rel
