"""Vocoder timing probe: wall time and per-kernel-class breakdown (engine profiler) for chunk batches."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine

FLOP_PER_FRAME = {10: 128.88e6, 30: 128.95e6, 90: 129.1e6, 160: 129.34e6, 1280: 132.78e6}
sd = W.make_random_weights(1234, wpe_rows=64)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=4, max_context=64,
           max_vocode_frames=int(os.environ.get("PROBE_FRAMES", "66000")))
g = torch.Generator().manual_seed(0)
cases = [(64, 10), (64, 30), (64, 90), (256, 160), (48, 1280)]
if os.environ.get("PROBE_CASES"):
    cases = [tuple(int(x) for x in c.split("x")) for c in os.environ["PROBE_CASES"].split(",")]
for (n, L) in cases:
    codes = torch.randint(0, 4096, (n * L,), generator=g).to("cuda", torch.int32)
    cu = list(range(0, (n + 1) * L, L))
    out = torch.empty((n * L * 320,), dtype=torch.float32, device="cuda")
    for _ in range(2):
        e.vocode(codes, cu, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        e.vocode(codes, cu, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = n * L * FLOP_PER_FRAME.get(L, 130e6)
    print(f"{n} x {L} frames: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  ({n * L / 75 / (ms / 1e3):9.0f} x real-time)", flush=True)
    if os.environ.get("PROBE_BREAKDOWN", "1") == "1":
        e.profile(2)
        e.vocode(codes, cu, out=out)
        rep = e.profile_report()
        e.profile(False)
        tot = sum(v["ms"] for v in rep.values())
        for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
            tf = v["flops"] / (v["ms"] / 1e3) / 1e12 if v["flops"] else 0.0
            print(f"     {k:20s} n={v['launches']:4d} {v['ms']:8.3f} ms ({100 * v['ms'] / tot:4.1f}%)  {tf:7.1f} TFLOP/s")
