"""Run a few steps of the hot path only (for ncu): python tools/profile_step.py [--model hubert_base] [--steps 2]
[--batch 32] [--seconds 10]. No CPU baseline, no e2e, no host copies — keep ncu captures short."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from s3prl_b200.upstream.expert import UpstreamExpert  # noqa: E402
from s3prl_b200.upstream.featurizer import weighted_sum  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="hubert_base")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--lanes", type=int, default=0)
args = ap.parse_args()

expert = UpstreamExpert(name=args.model, seed=0).to("cuda")
expert.lanes = args.lanes
g = torch.Generator().manual_seed(0)
wavs = [torch.randn(int(args.seconds * 16000), generator=g).cuda() for _ in range(args.batch)]
w = torch.softmax(torch.zeros(expert.num_layers + 1, device="cuda"), -1)
with torch.no_grad():
    for _ in range(args.warmup):
        weighted_sum(expert(wavs)["hidden_states"], w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        weighted_sum(expert(wavs)["hidden_states"], w)
    e1.record()
    torch.cuda.synchronize()
print(f"{args.model}: {e0.elapsed_time(e1) / args.steps:.3f} ms/step (not a bench number when run under ncu)")

# optional per-category device-time breakdown (CUDA events around every launch)
import ctypes as C  # noqa: E402

from s3prl_b200 import lib as s3lib  # noqa: E402

native = expert._native
s3lib.check(native.lib.s3b_profile_enable(native.handle, 1))
with torch.no_grad():
    for _ in range(args.steps):
        weighted_sum(expert(wavs)["hidden_states"], w)
ms5, fl5, ln5 = (C.c_double * 5)(), (C.c_double * 5)(), (C.c_int64 * 5)()
s3lib.check(native.lib.s3b_profile_read(native.handle, ms5, fl5, ln5, 1))
names = ["gemm", "attention", "conv0", "layernorm", "misc"]
print("breakdown ms/step:", {n: round(ms5[i] / args.steps, 3) for i, n in enumerate(names)},
      "launches/step:", sum(ln5) // args.steps)
