"""Diagnostic (not a test): per-layer relative error of the data2vec-style positional encoders vs the CPU oracle, printed
for every hidden state so that a failure can be localised (layer 0 = x + pos_conv blocks + LayerNorm)."""
import dataclasses
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))

import upstream_oracle as O  # noqa: E402
from s3prl_b200.upstream.configs import ARCHS  # noqa: E402
from s3prl_b200.upstream.expert import UpstreamExpert  # noqa: E402
from s3prl_b200.upstream.weights import fabricate_state_dict  # noqa: E402


def main():
    base = ARCHS["data2vec_base_960"]
    for depth, conv_pos, dim, lens in [(5, 95, 768, [24000, 17777, 1000]), (2, 95, 768, [24000, 17777]),
                                       (4, 96, 768, [24000, 17777]), (5, 95, 1024, [24000, 17777]),
                                       (3, 9, 768, [24000, 24000])]:
        cfg = dataclasses.replace(base, encoder_layers=2, pos_conv_depth=depth, conv_pos=conv_pos, encoder_embed_dim=dim,
                                  encoder_attention_heads=dim // 64, encoder_ffn_embed_dim=4 * dim)
        sd = fabricate_state_dict(cfg, seed=3)
        g = torch.Generator().manual_seed(77)
        wavs = [torch.randn(n, generator=g) for n in lens]
        with torch.no_grad():
            ref, _ = O.upstream_forward(wavs, sd, cfg)
        try:
            e = UpstreamExpert(arch=cfg, state_dict=sd, name="data2vec_variant").to("cuda")
            hs = e([w.cuda() for w in wavs])["hidden_states"]
            errs = [((h.cpu().double() - r.double()).norm() / r.double().norm()).item() for h, r in zip(hs, ref)]
            valid = e.valid_frames(lens)
            # error restricted to the valid frames of layer 0 (padded frames are computed too, but tell them apart)
            v0 = max(((hs[0][b, : valid[b]].cpu().double() - ref[0][b, : valid[b]].double()).norm()
                      / ref[0][b, : valid[b]].double().norm()).item() for b in range(len(lens)))
            print(f"depth={depth} conv_pos={conv_pos} k={cfg.pos_conv_kernel} D={dim} lens={lens}: "
                  f"rel err per hidden state {['%.2e' % x for x in errs]}  layer0 valid-only {v0:.2e}", flush=True)
            del e
        except Exception as ex:  # noqa: BLE001
            print(f"depth={depth} conv_pos={conv_pos} D={dim}: FAILED {type(ex).__name__}: {ex}", flush=True)


if __name__ == "__main__":
    main()
