"""Small end-to-end target for compute-sanitizer: both precisions, decode (kernel-per-op, graphs off via env, and the
fused kernel), ragged vocoder batch incl. one chunk long enough for the tensor-core attention, sampler."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine, Sampling

sd = W.make_random_weights(1234, wpe_rows=64)
rng = np.random.RandomState(0)
for precision in ("fp32", "bf16"):
    for fused in ("0", "1"):
        if precision == "fp32" and fused == "1":
            continue
        os.environ["LLMVOX_B200_FUSED"] = fused
        e = Engine(sd, device=0, precision=precision, max_sessions=8, max_context=48, max_vocode_frames=700)
        slots = [0, 3, 5]
        e.open(slots)
        e.feed_text(slots, [rng.randint(3, 259, size=k).tolist() for k in (20, 0, 7)])
        e.decode_steps(slots, 18)
        e.decode_step_logits(slots, sampling=Sampling(greedy=False, top_k=5, temperature=0.9, seed=1))
        codes = e.gather_code_ranges(slots, [0, 2, 4], [10, 5, 12])
        g = torch.Generator().manual_seed(1)
        extra = torch.randint(0, 4096, (300,), generator=g).to("cuda", torch.int32)
        allc = torch.cat([codes, extra]).contiguous()
        pcm = e.vocode(allc, [0, 10, 15, 27, 327])
        torch.cuda.synchronize()
        assert torch.isfinite(pcm).all()
        e.close()
        print(precision, "fused" if fused == "1" else "plain", "ok", flush=True)
