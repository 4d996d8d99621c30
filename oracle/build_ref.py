"""TEST INFRASTRUCTURE — recipe that installs the UNMODIFIED reference (s3prl @ /root/reference) into oracle/_ref.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: uses the travelled copy)

oracle/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so it travels with
the repo snapshot to the GPU box, where `bench.py --impl reference` times the reference's own
`s3prl.upstream.hubert.expert.UpstreamExpert` + `s3prl.upstream.interfaces.Featurizer` on the box's host cores
(cpu_baseline.kind = "reference") and `python -m s3prl_b200.run_downstream` imports the reference's Runner (config 5).

Steps: `pip install --no-index --no-build-isolation --no-deps --target oracle/_ref` from a scratch copy of the tree
(/root/reference is read-only and the build writes egg-info), then the data files of the one downstream recipe the
bench uses (downstream/ctc: yaml configs + vocabularies), which the wheel does not package.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REFERENCE = Path("/root/reference")
TARGET = ROOT / "oracle" / "_ref"


def is_current() -> bool:
    return (TARGET / "s3prl" / "upstream" / "hubert" / "expert.py").exists() and (
        TARGET / "s3prl" / "downstream" / "ctc" / "librispeech.yaml"
    ).exists()


def build(force: bool = False) -> Path | None:
    if not REFERENCE.exists():
        return TARGET if is_current() else None
    if is_current() and not force:
        return TARGET
    if TARGET.exists():
        shutil.rmtree(TARGET)
    TARGET.mkdir(parents=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "src"
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "*.pyc", "__pycache__", "result", "data"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(src)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"pip install of the reference failed:\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    ctc_src = REFERENCE / "s3prl" / "downstream" / "ctc"
    ctc_dst = TARGET / "s3prl" / "downstream" / "ctc"
    for f in ctc_src.glob("*.yaml"):
        shutil.copy2(f, ctc_dst / f.name)
    for d in ("vocab",):
        if (ctc_src / d).exists():
            shutil.copytree(ctc_src / d, ctc_dst / d, dirs_exist_ok=True)
    return TARGET


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
