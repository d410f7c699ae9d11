"""Small end-to-end target for compute-sanitizer (memcheck / racecheck): the three precisions, decode on the
cluster-resident kernel and on the kernel-per-op chain, ragged vocoder batch incl. one chunk long enough for the
tensor-core attention, sampler.

    compute-sanitizer --tool memcheck  python scripts/sanitizer_target.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_target.py cluster      # DSMEM exchange protocol only"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import _lib
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine, Sampling

only_cluster = len(sys.argv) > 1 and sys.argv[1] == "cluster"
sd = W.make_random_weights(1234, wpe_rows=64)
rng = np.random.RandomState(0)
for precision in ("bf16", "exact") if only_cluster else ("fp32", "bf16", "exact"):
    e = Engine(sd, device=0, precision=precision, max_sessions=24, max_context=48, max_vocode_frames=700)
    slots = [0, 3, 5] + list(range(6, 23))              # 20 sessions: two clusters, the second partly filled
    e.open(slots)
    e.feed_text(slots, [rng.randint(3, 259, size=int(k)).tolist() for k in rng.randint(0, 30, size=len(slots))])
    if precision != "fp32":
        e.decode_steps(slots, 6, path=_lib.PATH_CLUSTER)
    if not only_cluster:
        e.decode_steps(slots, 3, path=_lib.PATH_PER_OP)
        e.decode_step_logits(slots, sampling=Sampling(greedy=False, top_k=5, temperature=0.9, seed=1))
        codes = e.gather_code_ranges(slots[:3], [0, 2, 4], [5, 3, 4])
        g = torch.Generator().manual_seed(1)
        extra = torch.randint(0, 4096, (300,), generator=g).to("cuda", torch.int32)
        allc = torch.cat([codes, extra]).contiguous()
        pcm = e.vocode(allc, [0, 5, 8, 12, 312])
        torch.cuda.synchronize()
        assert torch.isfinite(pcm).all()
    torch.cuda.synchronize()
    e.close()
    print(precision, "ok", flush=True)
