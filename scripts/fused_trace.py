"""Phase-by-phase clock trace of the fused decode kernel (CTA 0, last iteration)."""
import os
import sys
os.environ["LLMVOX_B200_FUSED"] = "1"

os.environ["LLMVOX_B200_TRACE"] = "1"
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine

sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision="bf16", max_sessions=128, max_context=256, max_vocode_frames=1024)
rng = np.random.RandomState(0)
names = []
for l in range(4):
    names += [f"L{l}.qkv", f"L{l}.attn", f"L{l}.proj", f"L{l}.ln2", f"L{l}.fc", f"L{l}.proj2", f"L{l}.ln"]
names += ["lm_head"]
for n in (1, 64):
    slots = list(range(n))
    e.open(slots)
    e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
    e.decode_steps(slots, 100)
    e.decode_steps(slots, 20)   # traced: last iteration of this launch (its pick + next assembly are not stamped)
    t = np.array(e.peek_trace(1 + 2 * len(names)), dtype=np.int64)
    d = np.diff(t) / 1.965e3     # us at 1965 MHz
    work, sync = d[0::2], d[1::2]
    print(f"n={n}: iteration {d.sum():.1f} us; work {work.sum():.1f} us, barriers {sync.sum():.1f} us")
    full = np.array(e.peek_trace(256), dtype=np.int64)
    q0 = t[2 * names.index("L3.qkv")]        # stamp before the L3.qkv work
    fine = (full[200:208] - q0) / 1.965e3
    print("   L3.qkv fine (us since phase start): tma issued %.2f | first full %.2f | commits done %.2f | tmem_full seen %.2f | dumped %.2f | cluster sync1 %.2f | epilogue %.2f | cluster sync2 %.2f" % tuple(fine))
    for k, nm in enumerate(names):
        print(f"   {nm:10s} work {work[k]:6.2f} us   barrier {sync[k]:6.2f} us")
