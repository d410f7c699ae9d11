"""ncu target: the decode launches of ONE bench step exactly as bench.py's roofline leg runs them (64 streams, one lane:
four cluster_decode_kernel launches of 10 / 30 / 90 / 70 iterations, each followed by its vocoder batch).  The first
step warms up; capture the second with `--launch-skip 4 -c 4 -k regex:cluster_decode`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import LaneRunner
sd = W.make_random_weights(bench.SEED, wpe_rows=256)
S = bench.STREAMS
e = Engine(sd, device=0, precision=os.environ.get("TARGET_PRECISION", "exact"), max_sessions=2 * S, max_batch=S, max_context=208, max_vocode_frames=S * 96, decode_lanes=1)
texts = bench.synthetic_text(S, 1000)
pcm = torch.empty((S * bench.TOKENS * 320,), dtype=torch.float32, device=e.device)
runner = LaneRunner(e, 1)
for g in (list(range(S)), list(range(S, 2 * S))):
    e.open(g); e.feed_text(g, texts)
    bench.device_step(e, runner, g, pcm)
    torch.cuda.synchronize()
print("done", e.session_length(S))
