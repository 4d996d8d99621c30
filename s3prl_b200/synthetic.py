"""Synthetic LibriSpeech-shaped data for BASELINE config 5 (``run_downstream.py -u hubert_base -d ctc``).

There is no corpus and no audio I/O in this image, so the launcher can replace the ONE function the
CTC downstream expert uses to obtain its dataloaders — ``s3prl.downstream.ctc.data.load_dataset``
(s3prl/downstream/ctc/data.py:73-86, bound into ``ctc/expert.py`` at import) — by ``load_dataset`` below. Everything
downstream of it is the reference's own code, unchanged: ``Runner.train`` consumes ``(wavs, labels, filenames)`` with
``wavs`` a tuple of float32 numpy arrays sorted by descending length (ctc/data.py:40-43, runner.py:286-293).

Host-side batch assembly follows ``collect_audio_batch`` (ctc/data.py:11-43): buckets of ``batch_size`` utterances of
similar length, halved when the first utterance exceeds 300 000 samples, descending length inside the batch; lengths
are uniform in [2 s, 16 s] with an 8 % tail up to 24 s (train-clean-100's longest utterances, the
ones the halving rule exists for), labels are random character sequences of about 14 tokens per second. Waveforms are drawn once per dataset (a pool) so that the loader is not the bottleneck.
"""
from __future__ import annotations

from functools import partial
from typing import List, Sequence

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

SAMPLE_RATE = 16000
HALF_BATCH_SIZE_WAV_LEN = 300000  # ctc/data.py:11


class SyntheticLibriSpeech(Dataset):
    """Indexable like ``LibriDataset`` (ctc/corpus/librispeech.py:22-60): item = a bucket of (utterance id, tokens)."""

    def __init__(self, n_utts: int, vocab_size: int, bucket_size: int, seed: int = 1337, min_s: float = 2.0,
                 max_s: float = 16.0, tail_s: float = 24.0, tail_frac: float = 0.08):
        rng = np.random.default_rng(seed)
        lens = rng.integers(int(min_s * SAMPLE_RATE), int(max_s * SAMPLE_RATE) + 1, size=n_utts)
        tail = rng.random(n_utts) < tail_frac
        lens = np.where(tail, rng.integers(int(max_s * SAMPLE_RATE), int(tail_s * SAMPLE_RATE) + 1, size=n_utts), lens)
        max_s = tail_s
        texts = [rng.integers(4, max(5, vocab_size), size=max(1, int(14 * n / SAMPLE_RATE))).astype(np.int64) for n in lens]
        # LibriDataset sorts by transcription length, descending (librispeech.py:44-45): long utterances bucket together
        order = sorted(range(n_utts), key=lambda i: len(texts[i]), reverse=True)
        self.lens = [int(lens[i]) for i in order]
        self.texts = [texts[i] for i in order]
        self.bucket_size = bucket_size
        # one shared noise pool: utterance i is a window of it (cheap, deterministic, N(0,1) like the reference's
        # pseudo data, s3prl/util/pseudo_data.py:70)
        self.pool = rng.standard_normal(int(max_s * SAMPLE_RATE) + n_utts, dtype=np.float32)

    def wav(self, i: int) -> np.ndarray:
        return self.pool[i : i + self.lens[i]]

    def __getitem__(self, index):
        if self.bucket_size > 1:
            index = min(len(self.lens) - self.bucket_size, index)
            return [(i, self.texts[i]) for i in range(index, index + self.bucket_size)]
        return index, self.texts[index]

    def __len__(self):
        return len(self.lens)


def collect_synthetic_batch(batch, split: str, dataset: SyntheticLibriSpeech,
                            half_batch_size_wav_len: int = HALF_BATCH_SIZE_WAV_LEN):
    """``collect_audio_batch`` with the audio reader replaced by the synthetic pool (ctc/data.py:11-43)."""
    if type(batch[0]) is not tuple:
        batch = batch[0]
    first_len = dataset.lens[batch[0][0]]
    if split == "train" and first_len > half_batch_size_wav_len and len(batch) > 1:
        batch = batch[: len(batch) // 2]
    items = [(dataset.lens[i], f"synthetic-{i:06d}", dataset.wav(i), txt) for i, txt in batch]
    items.sort(key=lambda x: x[0], reverse=True)
    _lens, files, feats, texts = zip(*items)
    return feats, texts, files


def load_dataset(split: str, tokenizer, corpus: dict):
    """Drop-in for ``s3prl.downstream.ctc.data.load_dataset`` (same signature, same loader structure)."""
    from torch.distributed import is_initialized

    corpus = dict(corpus)
    corpus.pop("num_workers", None)
    batch_size = int(corpus.get("batch_size", 32))
    bucketing = bool(corpus.get("bucketing", True))
    vocab = int(getattr(tokenizer, "vocab_size", 32))
    if split == "train":
        dataset = SyntheticLibriSpeech(2048, vocab, batch_size if bucketing else 1, seed=1337)
        loader_bs = 1 if bucketing else batch_size
        sampler = DistributedSampler(dataset) if is_initialized() else None
        return DataLoader(dataset, batch_size=loader_bs, shuffle=(sampler is None), sampler=sampler,
                          collate_fn=partial(collect_synthetic_batch, split=split, dataset=dataset), num_workers=0)
    dataset = SyntheticLibriSpeech(16, vocab, 1, seed=4242 + len(split))
    return DataLoader(dataset, batch_size=1, shuffle=False,
                      collate_fn=partial(collect_synthetic_batch, split=split, dataset=dataset), num_workers=0)


def install(ctc_expert_module) -> None:
    """Rebind the name ``load_dataset`` inside ``s3prl.downstream.ctc.expert`` (imported there at module import,
    ctc/expert.py:14)."""
    ctc_expert_module.load_dataset = load_dataset


# ------------------------------------------------------------------------------------------------
# per-stage device timing of the unmodified Runner loop (config 5 measurement)
# ------------------------------------------------------------------------------------------------
class StageTimer:
    """CUDA-event timers around ``module.forward`` of the three models the Runner calls every step
    (runner.py:295-311): upstream, featurizer, downstream. Steps are delimited by upstream calls; whatever lies between
    the end of the downstream forward and the next upstream call (backward, optimizer, dataloader, H2D) is "rest"."""

    def __init__(self, skip: int = 3):
        self.skip = skip
        self.events = {"upstream": [], "featurizer": [], "downstream": []}
        self.step_marks: List[torch.cuda.Event] = []

    def wrap(self, cls, name: str):
        timer = self
        orig = cls.forward

        def forward(self_, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if name == "upstream":
                timer.step_marks.append(e0)
            out = orig(self_, *a, **k)
            e1.record()
            timer.events[name].append((e0, e1))
            return out

        cls.forward = forward

    def summary(self) -> dict:
        torch.cuda.synchronize()
        out = {}
        for name, evs in self.events.items():
            ts = [a.elapsed_time(b) for a, b in evs[self.skip :]]
            if ts:
                out[f"{name}_ms"] = sum(ts) / len(ts)
        marks = self.step_marks[self.skip :]
        if len(marks) > 1:
            total = marks[0].elapsed_time(marks[-1]) / (len(marks) - 1)
            out["step_ms"] = total
            out["rest_ms"] = total - sum(out.get(f"{n}_ms", 0.0) for n in self.events)
            out["steps_timed"] = len(marks) - 1
        return out
