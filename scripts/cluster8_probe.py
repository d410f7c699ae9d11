"""us per decode iteration of the two cuts of the cluster-resident kernel (16-CTA / 8-CTA clusters) and of the
kernel-per-op chain over batch sizes, greedy, PROBE_PRECISION = bf16 | exact; direct engine calls on one stream (no
LaneRunner policy)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import _lib, weights as W
from llmvox_b200.engine import Engine

N = int(os.environ.get("PROBE_N", "256"))
T0 = int(os.environ.get("PROBE_T0", "20"))
ITERS = int(os.environ.get("PROBE_ITERS", "100"))
sd = W.make_random_weights(1234, wpe_rows=512)
rng = np.random.RandomState(0)
PRECISION = os.environ.get("PROBE_PRECISION", "bf16")
e = Engine(sd, device=0, precision=PRECISION, max_sessions=N, max_batch=N, max_context=T0 + 2 * ITERS + 32, max_vocode_frames=256, decode_lanes=4)
print("cluster capacity (sessions per wave: 16-CTA, 8-CTA):", e.cluster_capacity(), flush=True)


def timed(n, path):
    slots = list(range(n))
    e.open(slots)
    e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
    e.decode_steps(slots, T0, path=path)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.decode_steps(slots, ITERS, path=path)
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / ITERS


names = {_lib.PATH_CLUSTER16: "16-CTA", _lib.PATH_CLUSTER8: "8-CTA", _lib.PATH_PER_OP: "per-op"}
for n in [int(x) for x in os.environ.get("PROBE_SIZES", "16,64,112,120,128,192,240,256").split(",")]:
    if n > N:
        continue
    row = []
    for path in (_lib.PATH_CLUSTER16, _lib.PATH_CLUSTER8, _lib.PATH_PER_OP):
        us = timed(n, path)
        row.append(f"{names[path]} {us:7.1f} us ({n / us * 1e6 / 75:8.0f} audio-s/s)")
    print(f"n={n:4d} T={T0}..{T0 + ITERS}: " + " | ".join(row), flush=True)
