"""Decode-iteration timing probe (CUDA events, no profiler): us per iteration for several batch / lane layouts."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import LaneRunner

ITERS = int(os.environ.get("PROBE_ITERS", "100"))   # LLMVOX_B200_FUSED=1 probes the fused kernel
sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=256, max_batch=256, max_context=256,
           max_vocode_frames=1024, decode_lanes=8)
rng = np.random.RandomState(0)


def timed(n, lanes, iters=ITERS, prefill=0):
    slots = list(range(n))
    e.open(slots)
    e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
    r = LaneRunner(e, lanes)
    r.sync_from_control()
    r.decode(slots, 10 + prefill)           # warm-up (+ graph capture) and context
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    r.decode(slots, iters)
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters, 1e6 * t_host / iters


if os.environ.get("PROBE_NCU") == "1":
    timed(64, 1, iters=20)
    sys.exit(0)
for n, lanes in [(64, 1), (16, 1), (1, 1), (256, 1), (64, 2), (64, 4), (64, 8), (256, 4), (256, 8)]:
    us, host = timed(n, lanes)
    print(f"n={n:4d} lanes={lanes}: {us:8.1f} us/iter (host enqueue {host:6.1f} us/iter)  T=20..{20 + ITERS}", flush=True)
us, host = timed(64, 1, prefill=100)
print(f"n=64 lanes=1 context 110..: {us:8.1f} us/iter")
e.profile(True)
timed(64, 1, iters=50)
rep = e.profile_report()
e.profile(False)
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:20s} launches {v['launches']:5d} avg {1e3 * v['ms'] / v['launches']:7.2f} us (incl. event overhead)")
