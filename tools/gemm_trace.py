"""Timeline of one CTA of the split-bf16 GEMM (clock64 stamps, cycles since kernel entry), for the shapes of one
transformer layer at a given token count:  python tools/gemm_trace.py [--tokens 2000]
  t1 prologue done  t2 after griddepcontrol.wait  t3 first k-block landed (MMA can start)
  t4 last MMA issued  t5 first tile accumulator complete  t6 first tile epilogue done
  t7/t8 same for the CTA's last tile  t9 teardown done
The library prints the line (S3B_GEMM_TRACE); the third of three back-to-back launches is the one traced."""
import argparse
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["S3B_GEMM_TRACE"] = "1"
import torch  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from s3prl_b200 import lib as s3lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=2000)
args = ap.parse_args()
lib = s3lib.load()
M = args.tokens
for name, N, K, res, gelu in [("qkv-like", 2304, 768, False, 0), ("out_proj", 768, 768, True, 0),
                              ("fc1", 3072, 768, False, 1), ("fc2", 768, 3072, True, 0)]:
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") * 0.02
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda")
    torch.cuda.synchronize()
    print(name, file=sys.stderr, end=": ", flush=True)
    s3lib.check(lib.s3b_linear_f32(C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                                   C.c_void_p(r.data_ptr() if res else None), M, N, K, gelu,
                                   C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
