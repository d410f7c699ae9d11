"""Cluster-resident decode kernel: quick parity check against the kernel-per-op path, us per iteration, phase trace."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import LaneRunner

sd = W.make_random_weights(1234, wpe_rows=512)
rng = np.random.RandomState(0)
os.environ["LLMVOX_B200_CLUSTER"] = "1"
if os.environ.get("PROBE_TRACE") == "1":
    os.environ["LLMVOX_B200_TRACE"] = "1"
N = int(os.environ.get("PROBE_N", "256"))
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=N, max_batch=N, max_context=512, max_vocode_frames=256, decode_lanes=8)
del os.environ["LLMVOX_B200_CLUSTER"]

if os.environ.get("PROBE_CHECK", "1") == "1":
    p = Engine(sd, device=0, precision=os.environ.get("PROBE_REF", "fp32"), max_sessions=32, max_context=64, max_vocode_frames=256)
    n = 20
    slots = list(range(n))
    texts = [rng.randint(3, 259, size=rng.randint(0, 40)).tolist() for _ in range(n)]
    for x in (e, p):
        x.open(slots)
        x.feed_text(slots, texts)
    worst = 0.0
    for t in range(20):
        e.decode_steps(slots, 1)
        codes = e.gather_codes(slots, t, 1).view(-1).contiguous()
        lf = e.peek_logits(n)
        lp, _ = p.decode_step_logits(slots, forced=codes)
        ok = bool((lf.argmax(dim=1).to(torch.int32) == codes).all())
        worst = max(worst, float((lf - lp).abs().max()))
        if t < 3 or not ok:
            print(f"step {t}: argmax ok {ok} max |dlogit| {float((lf - lp).abs().max()):.4g}", flush=True)
    print(f"parity: worst |dlogit| over 20 steps = {worst:.4g}", flush=True)
    p.close()


FORCED = {"16": 3, "8": 4}.get(os.environ.get("PROBE_CUT", ""))   # _lib.PATH_CLUSTER16 / PATH_CLUSTER8: bypass LaneRunner's policy


def timed(n, lanes, iters=100, prefill=0):
    slots = list(range(n))
    e.open(slots)
    e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
    r = LaneRunner(e, lanes)
    r.sync_from_control()
    run = (lambda k: e.decode_steps(slots, k, path=FORCED)) if FORCED else (lambda k: r.decode(slots, k))
    run(10 + prefill)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(iters)
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters


for n, lanes in [(16, 1), (1, 1), (64, 1), (64, 4), (128, 1), (128, 8), (256, 1)]:
    if n > N or os.environ.get("PROBE_QUICK") == "1":
        continue
    print(f"n={n:4d} lanes={lanes}: {timed(n, lanes):8.1f} us/iter  T=20..120", flush=True)
if os.environ.get("PROBE_QUICK") != "1":
    print(f"n={min(N, 64)} lanes=1 context 110..210: {timed(min(N, 64), 1, prefill=100):8.1f} us/iter", flush=True)
if os.environ.get("PROBE_TRACE") == "1":
    timed(int(os.environ.get("PROBE_TRACE_N", "16")), 1, iters=int(os.environ.get("PROBE_TRACE_ITERS", "50")),
          prefill=int(os.environ.get("PROBE_TRACE_PREFILL", "50")))
    tr = e.peek_trace(250) if hasattr(e, "peek_trace") else None
    if tr is not None:
        t0 = tr[0]
        names = ["start", "assemble"]
        for l in range(4):
            names += [f"L{l}." + x for x in ["ln1.sent", "ln1.xchg", "ln1.norm+signal", "qkv.acc", "qkv.xchg", "attn.done",
                                             "y.xchg+signal", "proj.acc", "ln2.sent", "ln2.xchg", "ln2.norm+signal",
                                             "fc.acc", "fc.epi+signal", "proj2.acc", "proj2.scatter", "proj2.xchg"]]
        names += ["lnf.sent", "lnf.xchg", "lnf.norm+signal", "lm.acc", "end"]
        prev = t0
        for nm, x in zip(names, tr):
            print(f"  {nm:18s} {(x - t0) / 1965.0:8.2f}  (+{(x - prev) / 1965.0:6.2f})")
            prev = x
