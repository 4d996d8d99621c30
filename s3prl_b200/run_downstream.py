"""Launcher: the reference's ``run_downstream.py`` with the B200-native upstreams injected into ``s3prl.hub``.

    python -m s3prl_b200.run_downstream -m train -u hubert_base -d ctc -c downstream/ctc/librispeech.yaml -n exp

is ``python run_downstream.py ...`` of the reference (s3prl/run_downstream.py:153-215) except that, before its
``main()`` runs, (1) stub modules are registered for optional third-party imports that this image lacks and that the
upstream path never touches, and (2) ``s3prl_b200.hub.install(s3prl.hub)`` replaces the hub entries
(``Runner._get_upstream`` does ``getattr(hub, args.upstream)``, s3prl/downstream/runner.py:141). The reference tree
is not modified; it only has to be importable (``pip install s3prl`` or ``PYTHONPATH=/path/to/s3prl``).
"""
from __future__ import annotations

import importlib
import sys
import types


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__s3prl_b200_stub__ = True
    sys.modules[name] = mod
    return mod


def _have(name: str) -> bool:
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def install_shims() -> list:
    """Register no-op stand-ins for imports the reference performs at module import time but that are absent or
    removed in this environment. Returns the names that were stubbed. None of them is used on the upstream path."""
    stubbed = []
    import torchaudio

    if not hasattr(torchaudio, "set_audio_backend"):  # removed in torchaudio >= 2.2 (run_downstream.py:157)
        torchaudio.set_audio_backend = lambda *a, **k: None
        stubbed.append("torchaudio.set_audio_backend")
    if not _have("torchaudio.sox_effects"):
        _stub("torchaudio.sox_effects", apply_effects_tensor=None, apply_effects_file=None)
        stubbed.append("torchaudio.sox_effects")
    if not _have("omegaconf"):
        class _Missing:  # noqa: N801
            def __init__(self, *a, **k):
                raise ImportError("omegaconf is not installed (only needed by data2vec / fairseq converters)")

        _stub("omegaconf", OmegaConf=_Missing, DictConfig=dict, II=lambda x: x, MISSING="???", open_dict=None,
              is_primitive_type=lambda *_: True)
        stubbed.append("omegaconf")
    if not _have("tensorboardX"):
        class SummaryWriter:  # minimal logger used by Runner (runner.py:267-268)
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def close(self):
                pass

        _stub("tensorboardX", SummaryWriter=SummaryWriter)
        stubbed.append("tensorboardX")
    if not _have("editdistance"):
        def _eval(a, b):
            prev = list(range(len(b) + 1))
            for i, x in enumerate(a, 1):
                cur = [i]
                for j, y in enumerate(b, 1):
                    cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
                prev = cur
            return prev[-1]

        _stub("editdistance", eval=_eval)
        stubbed.append("editdistance")
    try:
        import huggingface_hub

        for missing in ("HfFolder", "Repository"):  # removed in recent huggingface_hub (runner.py:27)
            if not hasattr(huggingface_hub, missing):
                setattr(huggingface_hub, missing, type(missing, (), {}))
                stubbed.append(f"huggingface_hub.{missing}")
    except Exception:
        _stub("huggingface_hub", HfApi=object, HfFolder=object, Repository=object)
        stubbed.append("huggingface_hub")
    return stubbed


def inject(featurizer: bool = True) -> list:
    """Import ``s3prl.hub`` (with shims) and replace its wav2vec2 / HuBERT / WavLM / fbank entries; with
    ``featurizer`` also replace the ``Featurizer`` class the Runner instantiates (``Runner._get_featurizer``,
    s3prl/downstream/runner.py:166-180 resolves the name imported at runner.py:24 from s3prl.upstream.interfaces) by
    the fused one, so that config 5 trains the layer weights through ``s3b_weighted_sum[_backward]`` instead of
    ``torch.stack`` + mul + sum over 13 x [B, T, D]."""
    install_shims()
    import s3prl.hub as ref_hub

    from . import hub as our_hub

    names = our_hub.install(ref_hub)
    if featurizer:
        import s3prl.upstream.interfaces as ref_interfaces

        from .upstream.featurizer import Featurizer

        ref_interfaces.Featurizer = Featurizer
        try:
            import s3prl.downstream.runner as ref_runner

            ref_runner.Featurizer = Featurizer
        except Exception as e:  # the runner's own optional imports (only needed for training)
            print(f"[s3prl_b200] s3prl.downstream.runner not importable yet ({e}); Featurizer injected into "
                  "s3prl.upstream.interfaces only", file=sys.stderr)
        names = names + ["Featurizer"]
    return names


def reject_unsupported(argv) -> None:
    """``-f/--upstream_trainable`` fine-tunes the upstream (run_downstream.py:65, runner.py:300-301). The B200
    upstreams are inference-only, so the run would silently train with a frozen upstream: refuse up front."""
    for a in argv:
        if a == "--upstream_trainable" or (a.startswith("-") and not a.startswith("--") and "f" in a[1:] and a[1:].isalpha()):
            raise SystemExit(
                "s3prl_b200: -f/--upstream_trainable is not supported (the B200-native upstreams are frozen, "
                "inference-only); run without it, as the SUPERB recipes do."
            )


def main():
    """Extra flags (consumed here, never seen by the reference's parser):
      --synthetic_data   replace the CTC expert's dataloaders by LibriSpeech-shaped synthetic batches (s3prl_b200/synthetic.py;
                         no corpus and no audio I/O exist in this image)
      --stage_timing     CUDA-event timers around the upstream / featurizer / downstream forwards; one JSON line
                         ("s3b_stage_timing") on stderr at exit
    Under torchrun the reference expects ``--local_rank`` (run_downstream.py:29,166-168); it is added from LOCAL_RANK."""
    import json
    import os

    argv = sys.argv[1:]
    synthetic = "--synthetic_data" in argv
    timing = "--stage_timing" in argv
    argv = [a for a in argv if a not in ("--synthetic_data", "--stage_timing")]
    reject_unsupported(argv)
    if "LOCAL_RANK" in os.environ and "--local_rank" not in argv and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        argv += ["--local_rank", os.environ["LOCAL_RANK"]]
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.argv = [sys.argv[0]] + argv
    if os.environ.get("S3B_NO_INJECT"):  # plumbing tests on a GPU-less host: the reference's own upstreams
        install_shims()
        names = []
    else:
        names = inject()
    print(f"[s3prl_b200] injected {len(names)} B200-native entries into s3prl.hub", file=sys.stderr)
    import s3prl

    # configs and vocabularies are addressed relative to the s3prl package directory (run_downstream.py:138-139,
    # ctc/librispeech.yaml:51): run from there, like `cd s3prl; python run_downstream.py ...`
    pkg_dir = os.path.dirname(os.path.abspath(s3prl.__file__))
    if "-o" in argv or "--override" in argv or True:
        for flag in ("-p", "--expdir"):
            if flag in argv:  # keep a user-given relative expdir relative to the original cwd
                i = argv.index(flag) + 1
                argv[i] = os.path.abspath(argv[i])
        sys.argv = [sys.argv[0]] + argv
    os.chdir(pkg_dir)
    if synthetic:
        import s3prl.downstream.ctc.expert as ctc_expert

        from . import synthetic as syn

        syn.install(ctc_expert)
        print("[s3prl_b200] CTC dataloaders replaced by synthetic LibriSpeech-shaped batches", file=sys.stderr)
    timer = None
    if timing:
        import atexit

        import s3prl.downstream.ctc.expert as ctc_expert

        from .synthetic import StageTimer
        from .upstream.expert import UpstreamExpert
        from .upstream.featurizer import Featurizer

        timer = StageTimer()
        timer.wrap(UpstreamExpert, "upstream")
        timer.wrap(Featurizer, "featurizer")
        timer.wrap(ctc_expert.DownstreamExpert, "downstream")

        def report():
            try:
                rec = timer.summary()
                rec["rank"] = int(os.environ.get("RANK", "0"))
                rec["world"] = int(os.environ.get("WORLD_SIZE", "1"))
                print("s3b_stage_timing " + json.dumps(rec), file=sys.stderr, flush=True)
            except Exception as e:  # never mask the run's own exit status
                print(f"s3b_stage_timing failed: {e}", file=sys.stderr)

        atexit.register(report)
    from s3prl import run_downstream

    run_downstream.main()


if __name__ == "__main__":
    main()
