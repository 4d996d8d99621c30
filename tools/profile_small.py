"""Launch the small kernel families once each for `ncu --set full` (profiles/r2_small_kernels_ncu.txt): LayerNorm,
weighted sum, WavLM gate (through a wavlm_base_plus forward), fbank, mel spectrogram.

    ncu --set full --clock-control none --import-source on \
        -k regex:"layernorm_kernel|weighted_sum_kernel|wavlm_gate|fbank|stft_mel|mel_cmvn|conv0_moments|wav_pack|posconv_combine" \
        -c 24 -f -o gpurun_out/prof_small python tools/profile_small.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from s3prl_b200 import hub  # noqa: E402
from s3prl_b200.upstream.featurizer import weighted_sum  # noqa: E402

g = torch.Generator().manual_seed(0)
wavs = [torch.randn(160000, generator=g).cuda() for _ in range(32)]
with torch.no_grad():
    e = hub.wavlm_base_plus().to("cuda")
    e.lanes = 1
    w = torch.softmax(torch.zeros(13, device="cuda"), -1)
    for _ in range(2):
        weighted_sum(e(wavs)["hidden_states"], w)
    del e
    fb, mel = hub.fbank().to("cuda"), hub.mel().to("cuda")
    short = [x[:16000].contiguous() for x in wavs[:4]]  # BASELINE C1: 4 x 1 s
    for _ in range(2):
        fb(short), mel(short)
        fb(wavs), mel(wavs)  # 32 x 10 s: the throughput shape
torch.cuda.synchronize()
print("ok")
