"""Per-kernel breakdown of decode iterations at large batch (config 4: 2048 sessions per GPU; also 256 / 512): engine
profiler (CUDA events around every launch), kernel-per-op chain."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import _lib, weights as W
from llmvox_b200.engine import Engine

N = int(os.environ.get("PROBE_N", "2048"))
T0 = int(os.environ.get("PROBE_T0", "300"))
STEPS = int(os.environ.get("PROBE_STEPS", "10"))
sd = W.make_random_weights(1234, wpe_rows=T0 + STEPS + 64)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=N, max_batch=N, max_context=T0 + 2 * STEPS + 32,
           max_vocode_frames=256, decode_lanes=1)
rng = np.random.RandomState(0)
slots = list(range(N))
e.open(slots)
e.feed_text(slots, [rng.randint(3, 259, size=50).tolist() for _ in slots])
PATH = {'per_op': _lib.PATH_PER_OP, 'cluster': _lib.PATH_CLUSTER, 'auto': _lib.PATH_AUTO}[os.environ.get('PROBE_PATH', 'per_op')]
e.decode_steps(slots, T0, path=PATH)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
e.decode_steps(slots, STEPS, path=PATH)
b.record()
torch.cuda.synchronize()
print(f"n={N} T={T0}..{T0 + STEPS}: {1e3 * a.elapsed_time(b) / STEPS:.1f} us / iteration (graphs + PDL)")
e.profile(2 if os.environ.get('PROBE_DETAIL') else True)
e.decode_steps(slots, STEPS, path=PATH)
rep = e.profile_report()
e.profile(False)
tot = sum(v["ms"] for v in rep.values())
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0.0
    gb = v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["bytes"] else 0.0
    print(f"   {k:18s} n={v['launches']:5d} {1e3 * v['ms'] / STEPS:9.1f} us/iter ({100 * v['ms'] / tot:4.1f}%) avg {1e3 * v['ms'] / v['launches']:7.1f} us  {tf:7.1f} TFLOP/s {gb:7.1f} GB/s")
