"""BASELINE config 4 shape: many concurrent sessions, 600-code utterances, end to end (host text ids -> PCM in host
memory), session-sharded over the ranks of a torchrun launch (no collective on the data path).

    python scripts/config4_scale.py --sessions 4096            # 1 GPU
    torchrun --nproc-per-node 2 ... scripts/config4_scale.py --sessions 4096
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.sharding import shard_indices
from llmvox_b200.streaming import BatchSynthesizer

ap = argparse.ArgumentParser()
ap.add_argument("--sessions", type=int, default=4096)
ap.add_argument("--tokens", type=int, default=600)
ap.add_argument("--wave", type=int, default=0, help="sessions decoded together per GPU (0 = all of the rank's sessions)")
ap.add_argument("--lanes", type=int, default=1)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mine = shard_indices(args.sessions, world, rank)
wave = args.wave or len(mine)
sd = W.make_random_weights(1234, wpe_rows=args.tokens + 8)
e = Engine(sd, device=local, precision="bf16", max_sessions=wave, max_batch=wave, max_context=args.tokens + 8,
           max_vocode_frames=65536, decode_lanes=args.lanes)
rng = np.random.RandomState(rank)
texts = {i: rng.randint(3, 259, size=args.tokens).tolist() for i in mine}


def run_once():
    codes_done = 0
    for w0 in range(0, len(mine), wave):
        ids = mine[w0:w0 + wave]
        bs = BatchSynthesizer(e, len(ids), 10, stop_on_eoa=False, lanes=args.lanes)
        bs.start([texts[i] for i in ids])
        for chunks in bs.run(args.tokens, flush_tail=True, copy=False):
            codes_done += sum(c.length for c in chunks)
    return codes_done


run_once()
best = None
for _ in range(args.reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    n = run_once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = float(t[0]) if best is None else min(best, float(t[0]))
    assert n == len(mine) * args.tokens
if rank == 0:
    audio = args.sessions * args.tokens / 75.0
    print(f"config4: {args.sessions} sessions x {args.tokens} codes on {world} GPU(s), wave {wave}/GPU, lanes {args.lanes}: "
          f"{best:.3f} s -> {audio / best:.0f} audio-s/s ({audio / best / world:.0f} per GPU), device memory {e.device_bytes / 2**30:.1f} GiB")
e.close()
if world > 1:
    dist.destroy_process_group()
