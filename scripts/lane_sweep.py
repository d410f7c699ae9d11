"""us per decode iteration for 64 sessions vs number of decode lanes (engine budget sized for that lane count)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import LaneRunner
sd = W.make_random_weights(1234, wpe_rows=256)
rng = np.random.RandomState(0)
n = int(os.environ.get("PROBE_N", "64"))
for lanes in (1, 2, 3, 4, 5, 6, 8):
    e = Engine(sd, device=0, precision="bf16", max_sessions=n, max_context=256, max_vocode_frames=256, decode_lanes=lanes)
    slots = list(range(n))
    e.open(slots); e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
    r = LaneRunner(e, lanes); r.sync_from_control(); r.decode(slots, 20); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r.decode(slots, 100); b.record(); torch.cuda.synchronize()
    print(f"n={n} lanes={lanes}: {1e3 * a.elapsed_time(b) / 100:7.1f} us/iter", flush=True)
    e.close()
