"""Timing probes of the f16q8 GEMM main loop (S3B_Q8_DEBUG variants, numerics intentionally wrong for 1-4)."""
import ctypes as C, os, sys, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
if len(sys.argv) == 1:
    for scheme, dbg in (("bf16x3", "0"), ("f16q8", "0"), ("f16q8", "1"), ("f16q8", "2"), ("f16q8", "3"), ("f16q8", "4")):
        env = dict(os.environ, S3B_GEMM_SCHEME=scheme, S3B_Q8_DEBUG=dbg)
        r = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True)
        print(f"scheme {scheme} debug {dbg}: {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)
else:
    from s3prl_b200 import lib
    L = lib.load()
    out = (C.c_float * 2)()
    res = []
    for (M, N, K, gelu) in ((15968, 2304, 768, 0), (15968, 768, 3072, 0), (15968, 3072, 768, 1), (1996, 768, 3072, 0)):
        for un in (128, 256):
            lib.check(L.s3b_gemm_bench(M, N, K, gelu, un, 30, out))
            res.append(f"{M}x{N}x{K}{'g' if gelu else ''}/un{un}={out[0]*1e3:.1f}us")
    print("  ".join(res))
