"""Small target for `ncu --set full`: 3 decode iterations of 64 sessions (kernel-per-op path, run with
LLMVOX_B200_NO_GRAPH=1 so every kernel is a plain launch) and one vocoder pass over 64 chunks of 160 codes.
tc_gemm_kernel launch order: 3 x 17 swap-mode decode GEMMs, then 37 normal-mode vocoder GEMMs (embed k7, 2 x 2 resnet
k3, attn qkv, attn proj, 2 x 2 resnet k3, 12 x (pw1 + GELU, pw2 + residual), head, iDFT)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine

sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision="bf16", max_sessions=64, max_context=256, max_vocode_frames=64 * 170)
rng = np.random.RandomState(0)
slots = list(range(64))
e.open(slots)
e.feed_text(slots, [rng.randint(3, 259, size=200).tolist() for _ in slots])
e.decode_steps(slots, 3)
g = torch.Generator().manual_seed(0)
codes = torch.randint(0, 4096, (64 * 160,), generator=g).to("cuda", torch.int32)
pcm = e.vocode(codes, list(range(0, 65 * 160, 160)))
torch.cuda.synchronize()
print("ok", float(pcm.abs().mean()))
