"""ncu target: one launch of the cluster-resident decode kernel (PROBE_N sessions, 20 iterations at context ~110).
PROBE_PRECISION = bf16 | exact; PROBE_CUT = 16 | 8 forces the cut (8-CTA clusters: bf16 only)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LLMVOX_B200_CLUSTER"] = "1"
from llmvox_b200 import _lib, weights as W
from llmvox_b200.engine import Engine
sd = W.make_random_weights(1234, wpe_rows=256)
n = int(os.environ.get("PROBE_N", "16"))
path = {"16": _lib.PATH_CLUSTER16, "8": _lib.PATH_CLUSTER8}.get(os.environ.get("PROBE_CUT", ""), _lib.PATH_CLUSTER)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=n, max_batch=n, max_context=256, max_vocode_frames=256)
rng = np.random.RandomState(0)
slots = list(range(n))
e.open(slots); e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
e.decode_steps(slots, 110, path=path)
torch.cuda.synchronize()
e.decode_steps(slots, 20, path=path)
torch.cuda.synchronize()
print("done", e.session_length(0))
