"""Teacher-forced logit error against the fp32 engine (itself pinned to the reference within 1e-4): the cluster-resident
decode kernel and the kernel-per-op bf16 path, same histories (the cluster kernel's greedy picks)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
sd = W.make_random_weights(1234, wpe_rows=64)
n, steps = int(os.environ.get("PROBE_N", "64")), 40
kw = dict(device=0, max_sessions=n, max_context=48, max_vocode_frames=256)
clus = Engine(sd, precision="bf16", **kw)
os.environ["LLMVOX_B200_CLUSTER"] = "0"
plain = Engine(sd, precision="bf16", **kw)
del os.environ["LLMVOX_B200_CLUSTER"]
ref = Engine(sd, precision="fp32", **kw)
rng = np.random.RandomState(n)
texts = [rng.randint(3, 259, size=rng.randint(0, 40)).tolist() for _ in range(n)]
slots = list(range(n))
for e in (clus, plain, ref):
    e.open(slots); e.feed_text(slots, texts)
wc = wp = wcp = 0.0
rc = rp = 0.0
for t in range(steps):
    clus.decode_steps(slots, 1)
    codes = clus.gather_codes(slots, t, 1).view(-1).contiguous()
    lc = clus.peek_logits(n)
    lp, _ = plain.decode_step_logits(slots, forced=codes)
    lr, _ = ref.decode_step_logits(slots, forced=codes)
    wc = max(wc, float((lc - lr).abs().max())); wp = max(wp, float((lp - lr).abs().max())); wcp = max(wcp, float((lc - lp).abs().max()))
    rc += float(((lc - lr) ** 2).mean()); rp += float(((lp - lr) ** 2).mean())
print(f"n={n} steps={steps}: max|cluster-fp32|={wc:.4g}  max|plain_bf16-fp32|={wp:.4g}  max|cluster-plain_bf16|={wcp:.4g}  "
      f"rms cluster {np.sqrt(rc / steps):.4g}  rms plain {np.sqrt(rp / steps):.4g}  logit std {float(lr.std()):.3g}")
