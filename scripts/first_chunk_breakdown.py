"""Where the first-chunk latency of a 64-stream batch goes: the phases of BatchSynthesizer's first round, each followed by
a device synchronisation (so the sum is an upper bound of the pipelined latency bench.py reports)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import BatchSynthesizer

B = int(os.environ.get("PROBE_B", "64"))
sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "exact"), max_sessions=B, max_batch=B, max_context=256,
           max_vocode_frames=B * 170, decode_lanes=4)
rng = np.random.RandomState(0)
texts = [rng.randint(3, 259, size=200).tolist() for _ in range(B)]
bs = BatchSynthesizer(e, B, 10, stop_on_eoa=False, lanes=4)
rows = []
for rep in range(12):
    torch.cuda.synchronize()
    t = [time.perf_counter()]
    bs.start(texts); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    bs.runner.decode(bs.slots, 10, bs.sampling); t.append(time.perf_counter())
    torch.cuda.synchronize(); t.append(time.perf_counter())
    ticket = bs._enqueue_emit([(i, 0, 10) for i in range(B)]); t.append(time.perf_counter())
    chunks = bs._finish_emit(ticket, False); t.append(time.perf_counter())
    rows.append(np.diff(t) * 1e3)
r = np.median(np.array(rows[2:]), axis=0)
names = ["start() host", "start() device", "decode(10) host", "decode(10) device", "gather + vocode + D2H enqueue (host)", "wait for PCM"]
for n, v in zip(names, r):
    print(f"{n:40s} {v:7.3f} ms")
print(f"{'sum':40s} {r.sum():7.3f} ms")
