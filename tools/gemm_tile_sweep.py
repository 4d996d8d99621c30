"""Tile-shape sweep of the production GEMM at the token counts of the sharded runs (N = 1, 2, 4, 8 GPUs x 1 or 2 lanes):
python tools/gemm_tile_sweep.py > gpurun_out/gemm_tile_sweep.txt. Output feeds pick_pair_umma_n (model.cu)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from s3prl_b200 import lib

L = lib.load()
out = (C.c_float * 2)()
shapes = {  # (N, K, gelu)
    "base qkv": (2304, 768, 0), "base out": (768, 768, 0), "base fc1": (3072, 768, 1), "base fc2": (768, 3072, 0),
    "large qkv": (3072, 1024, 0), "large out": (1024, 1024, 0), "large fc1": (4096, 1024, 1), "large fc2": (1024, 4096, 0),
}
print("M      shape        un128_us  un256_us  default(un)")
for utts, T in ((2, 499), (4, 499), (8, 499), (16, 499), (32, 499), (1, 999), (2, 999), (4, 999), (8, 999), (16, 999)):
    M = utts * T
    for name, (N, K, gelu) in shapes.items():
        if (T == 499) != name.startswith("base"):
            continue
        res = []
        for un in (128, 256, 0):
            lib.check(L.s3b_gemm_bench(M, N, K, gelu, un, 30, out))
            res.append((out[0] * 1e3, int(out[1])))
        flops = 6.0 * M * N * K
        print(f"{M:6d} {name:10s} {res[0][0]:9.2f} {res[1][0]:9.2f} {res[2][0]:9.2f} ({res[2][1]})   ideal {flops / 1344e12 * 1e6:7.2f} us")
