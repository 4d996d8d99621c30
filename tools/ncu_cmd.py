"""Print the two ncu command lines used for profiles/ (see profiles/README.md), with the -s/-c launch offsets that
select layer 0 of the SECOND (warm) step of tools/profile_step.py for a given architecture:

    python tools/ncu_cmd.py [--model hubert_base] [--tag r2a]

Kernels matched by the regex, in launch order per step (post-LN or pre-LN alike):
    conv0_apply (1) | conv1..6 GEMMs (6) | post_extract_proj GEMM (1) | pos_conv GEMM (1) | posconv_combine (1) |
    per layer: QKV GEMM, attention, out_proj GEMM, fc1 GEMM, fc2 GEMM (5)
`layer_norm` extractors (large_ll60k, hubert_large, wavlm_large, unispeech_sat_large) use conv0_apply_ln instead of
conv0_apply_gn — the regex `conv0_apply` matches both."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from s3prl_b200.upstream.configs import get_arch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="hubert_base")
ap.add_argument("--tag", default="rXy")
args = ap.parse_args()
cfg = get_arch(args.model)
front = 1 + 6 + 1 + 1 + 1
per_step = front + 5 * cfg.encoder_layers
skip = per_step + 7  # second step: skip conv0_apply + conv1..6, keep proj, pos_conv, combine and layer 0
regex = "gemm2|attention_kernel|conv0_apply|posconv_combine"
step = f"python tools/profile_step.py --model {args.model} --steps 1 --warmup 1"
print("# 1) launch list + DRAM bytes per launch (tools/ncu_summary.py launches / tools/ncu_traffic.py read the CSV)")
print(f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv "
      f"--log-file gpurun_out/launches_{args.tag}.csv {step}")
print(f"# 2) full sections of proj, pos_conv, combine, QKV, attention, out_proj, fc1, fc2 ({per_step} matching launches per step)")
print(f'ncu --set full --clock-control none --import-source on -k regex:"{regex}" -s {skip} -c 8 -f '
      f"-o gpurun_out/prof_{args.tag} {step}")
print("# then: python tools/ncu_summary.py report gpurun_out/prof_%s.ncu-rep > profiles/%s_kernels_ncu.txt" % (args.tag, args.tag))
