#!/usr/bin/env python
"""bench.py — upstream frames/s of the B200 hot path (BASELINE.json metric) and the reference-CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the oracle port on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # N > 1: one rank per GPU over NCCL

Workload (BASELINE.json configs[1]): hubert_base (12L, 768d), global batch 32 x 10 s of synthetic 16 kHz audio
(seeded N(0,1) samples, fabricated random-init checkpoint — no network for data or weights). A "step" is one
pass of the hot path over the batch: UpstreamExpert(wavs) -> 13 hidden states [B, 499, 768] fp32 materialised,
then the Featurizer weighted sum; with N > 1 the batch is sharded by utterance (strong scaling, Lmax shared) and
the step ends with ONE NCCL all-gather of the weighted-sum features.

One JSON line on stdout (rank 0). Keys beyond the base contract: "roofline" (tcgen05 GEMM kernel, algorithmic
FLOP/s from CUDA events around every GEMM launch in a profiled pass of the same steps), "cpu_baseline"
(oracle port on a bounded sample, rank 0, N=1), "e2e" (same metric through s3b_forward_host: pinned host
waveforms in, all hidden states back to pinned host memory, copies inside the timed region), "clocks".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL = "hubert_base"
GLOBAL_BATCH = 32
SECONDS = 10
SAMPLE_RATE = 16000
METRIC = "upstream frames/sec (16 kHz) hubert_base @ batch=32x10 s"


def algorithmic_flops_per_utt(cfg, L):
    """conv + proj + posconv + linear + attn on the true frame count (multiply-add = 2)."""
    convs = [(512, 10, 5)] + [(512, 3, 2)] * 4 + [(512, 2, 2)] * 2
    n, cin, conv = L, 1, 0.0
    for dim, k, s in convs:
        n = (n - k) // s + 1
        conv += 2.0 * cin * dim * k * n
        cin = dim
    T, D, F, NL = n, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_layers
    proj = 2.0 * 512 * D * T
    pos = 2.0 * D * (D // cfg.conv_pos_groups) * cfg.conv_pos * T
    lin = NL * (8.0 * D * D + 4.0 * D * F) * T
    attn = NL * 4.0 * T * T * D
    return conv + proj + pos + lin + attn, T


def gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch (average over the launches of one step) from the
    committed ncu capture of `tools/profile_step.py` (profiles/r1_traffic.json, see profiles/README.md); None if absent."""
    try:
        t = json.loads((ROOT / "profiles" / "r1_traffic.json").read_text())
        fam = [v for k, v in t.items() if k.startswith("gemm")]
        n = sum(v["launches"] for v in fam)
        return sum(v["dram_bytes_per_launch"] * v["launches"] for v in fam) / n if n else None
    except Exception:
        return None


def seeded_wav(idx: int, n: int):
    import torch

    g = torch.Generator().manual_seed(1000 + idx)
    return torch.randn(n, generator=g)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def stop(self):
        """Summary of the samples taken between mark_begin() and now (the timed + profiled passes)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < getattr(self, "t0", 0.0) or ts > t1 + 0.03:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1])), power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def usable_cores() -> int:
    """Host cores this process may really use: min(affinity mask, cgroup v2/v1 CPU quota). On the GPU boxes
    os.cpu_count() reports 128 while the container is capped at 16 CPUs; 128 threads would thrash."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            p = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    return n


def cpu_oracle_throughput(n_utts: int, steps: int, warmup: int):
    """Time the oracle port (torch-CPU restatement of the reference forward) on `n_utts` x 10 s of the workload."""
    import torch

    sys.path.insert(0, str(ROOT / "oracle"))
    import upstream_oracle as O
    from s3prl_b200.upstream.configs import get_arch
    from s3prl_b200.upstream.weights import fabricate_state_dict

    cores = usable_cores()
    torch.set_num_threads(cores)
    cfg = get_arch(MODEL)
    sd = fabricate_state_dict(cfg, seed=0)
    wavs = [seeded_wav(i, SECONDS * SAMPLE_RATE) for i in range(n_utts)]
    w = torch.zeros(cfg.encoder_layers + 1)
    times = []
    frames = 0
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            hs, _ = O.upstream_forward(wavs, sd, cfg)
            feat = O.weighted_sum(hs, w)
            dt = time.perf_counter() - t0
            frames = feat.shape[0] * feat.shape[1]
            if it >= warmup:
                times.append(dt)
    total = sum(times)
    return frames * len(times) / total, 1e3 * total / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    n_utts = 4 if cores < 48 else 8
    fps, ms, cores = cpu_oracle_throughput(n_utts, args.steps, args.warmup)
    sample = f"{n_utts} of the 32 utterances (x 10 s) per step; oracle port of the reference forward, torch-CPU fp32, {cores} threads"
    line = {
        "impl": "reference",
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} forward + featurizer weighted sum, global batch {GLOBAL_BATCH} x {SECONDS} s @16 kHz",
                   "sample": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from s3prl_b200 import lib as s3lib
    from s3prl_b200.upstream.expert import UpstreamExpert
    from s3prl_b200.upstream.featurizer import weighted_sum

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU); see module docstring")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    L = SECONDS * SAMPLE_RATE
    assert GLOBAL_BATCH % world == 0
    per = GLOBAL_BATCH // world
    my_ids = list(range(rank * per, (rank + 1) * per))
    expert = UpstreamExpert(name=MODEL, seed=0).to(device)
    expert.global_max_len = L  # padding / GroupNorm statistics identical to the un-sharded batch
    cfg = expert.arch
    wavs_host = [seeded_wav(i, L).pin_memory() for i in my_ids]
    wavs = [w.to(device) for w in wavs_host]
    NLp1, D = cfg.encoder_layers + 1, cfg.encoder_embed_dim
    fw = torch.softmax(torch.zeros(NLp1, device=device), -1)
    T = expert.num_frames(L)
    gathered = torch.empty(GLOBAL_BATCH, T, D, device=device) if world > 1 else None

    def step():
        res = expert(wavs)
        feat = weighted_sum(res["hidden_states"], fw)
        if world > 1:
            dist.all_gather_into_tensor(gathered, feat)
        return feat

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()  # nvidia-smi needs ~100 ms to come up: start it before the warm-up
        for _ in range(max(args.warmup, 3)):
            step()
        # ---- timed region: device-resident inputs ----------------------------------------------------------
        native = expert._native
        launches0 = native.lib.s3b_launch_count(native.handle)
        barrier()
        if sampler:
            sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        ms_total = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        ms_step = ms_total.item() / args.steps
        launches = (native.lib.s3b_launch_count(native.handle) - launches0) + args.steps  # + weighted-sum kernel
        frames_step = GLOBAL_BATCH * T
        value = frames_step / (ms_step * 1e-3)

        # ---- profiled pass: CUDA events around every launch, same steps (GEMM roofline) ------------------------
        import ctypes as C

        s3lib.check(native.lib.s3b_profile_enable(native.handle, 1))
        barrier()
        for _ in range(args.steps):
            step()
        ms5, fl5, ln5 = (C.c_double * 5)(), (C.c_double * 5)(), (C.c_int64 * 5)()
        s3lib.check(native.lib.s3b_profile_read(native.handle, ms5, fl5, ln5, 1))
        s3lib.check(native.lib.s3b_profile_enable(native.handle, 0))
        clocks = sampler.stop() if sampler else None  # samples of the timed + profiled passes (same kernels, same load)
        cat = ["gemm_tcgen05", "attention_tcgen05", "conv0_norm_gelu", "layernorm", "misc"]
        breakdown = {c: {"ms_per_step": ms5[i] / args.steps, "launches_per_step": ln5[i] // args.steps,
                         "alg_tflop_per_step": fl5[i] / args.steps / 1e12} for i, c in enumerate(cat)}
        gemm_tflops = (fl5[0] / 1e12) / (ms5[0] * 1e-3) if ms5[0] > 0 else 0.0

        # ---- e2e: host buffers through s3b_forward_host (H2D + D2H inside the timed region) -------------------
        for _ in range(2):
            expert.forward_host(wavs_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_host = expert.forward_host(wavs_host)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_value = frames_step * args.steps / dt.item()
        h2d = sum(w.numel() * 4 for w in wavs_host)
        d2h = out_host.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    flops_utt, _ = algorithmic_flops_per_utt(cfg, L)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 (bf16 hi+lo split operands x3 MMAs, fp32 TMEM accumulate)", "data": "synthetic",
        "config": {
            "workload": f"{MODEL} forward + featurizer weighted sum, global batch {GLOBAL_BATCH} x {SECONDS} s @16 kHz"
                        + (f", sharded {per}/GPU + 1 NCCL all-gather" if world > 1 else ""),
            "frames_per_step": frames_step, "l2": "activations and outputs (0.64 GB of hidden states per step) exceed the 126 MB L2",
            "alg_tflop_per_step": flops_utt * GLOBAL_BATCH / 1e12,
        },
        "whole_step_tflops": flops_utt * GLOBAL_BATCH / world / 1e12 / (ms_step * 1e-3) * world,
        "roofline": {
            "kernel": "gemm2_bf16x3_kernel (tcgen05 cta_group::2, all GEMM launches of a step)", "bound": "tensor",
            "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf,
            "traffic": gemm_traffic(), "peak_source": peak_src,
            "mma_pipe_tflops": 3.0 * gemm_tflops,
            "note": "achieved = algorithmic FLOPs (1 MMA per product; the tensor pipe executes 3 bf16 MMAs per product = mma_pipe_tflops) / CUDA-event time per launch, rank 0; traffic = bytes per launch (ncu, profiles/)",
        },
        "kernel_breakdown": breakdown,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "host_chunks": int(os.environ.get("S3B_HOST_CHUNKS", 2 if GLOBAL_BATCH // world >= 16 else 1))},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        fps, ms, cores = cpu_oracle_throughput(4, 3, 1)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "4 of the 32 utterances (x 10 s), 1 warm-up + 3 timed passes of the oracle port (torch-CPU fp32)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
