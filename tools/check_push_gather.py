"""2+ GPU check of the fused weighted-sum + peer-memory push all-gather against weighted_sum + NCCL all-gather
(bit-exact), several steps deep so that slot reuse and the flag protocol are exercised:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_push_gather.py"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from s3prl_b200.parallel import FeatureGatherer
from s3prl_b200.upstream.featurizer import weighted_sum

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
NL, B, T, D = 13, 3, 499, 768
g = FeatureGatherer((B, T, D), dev, mode="push")
print(f"rank {rank}: gather mode {g.mode}", flush=True)
ref = torch.empty(world * B, T, D, device=dev)
ok = True
outs = []
for step in range(9):
    gen = torch.Generator(device=dev).manual_seed(1000 * step + rank)
    hs = torch.randn(NL, B, T, D, device=dev, generator=gen)
    w = torch.softmax(torch.randn(NL, device=dev, generator=gen), -1)
    dist.broadcast(w, 0)
    local_feat = weighted_sum([hs[i] for i in range(NL)], w)
    dist.all_gather_into_tensor(ref, local_feat)
    if step % 3 == 1:  # let ranks drift: the protocol must hold with one step of slack
        torch.cuda._sleep(int(2e8) * (rank + 1))
    got = g.weighted_sum_gather([hs[i] for i in range(NL)], w)
    outs.append((got, ref.clone()))
    if len(outs) >= 2:  # the tensor of the PREVIOUS call is complete after this call
        a, b = outs[-2]
        torch.cuda.synchronize()
        same = torch.equal(a, b)
        ok &= same
        if not same:
            print(f"rank {rank} step {step - 1}: MISMATCH max {(a - b).abs().max().item()}", flush=True)
g.finish()
torch.cuda.synchronize()
ok &= torch.equal(*outs[-1])
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PUSH_GATHER_OK" if flag.item() == 1 else "PUSH_GATHER_FAILED", flush=True)
g.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
