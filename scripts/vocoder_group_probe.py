"""Config 3 (1024 chunks x 1280 frames) through lvx_vocode at different launch-group sizes (max_vocode_frames)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
sd = W.make_random_weights(1234, wpe_rows=32)
n, L = 1024, 1280
g = torch.Generator().manual_seed(3)
codes = torch.randint(0, 4096, (n * L,), generator=g).to("cuda", torch.int32)
cu = list(range(0, (n + 1) * L, L))
out = torch.empty((n * L * 320,), dtype=torch.float32, device="cuda")
for group in [int(x) for x in os.environ.get("GROUPS", "48,64,96,128").split(",")]:
    e = Engine(sd, device=0, precision="bf16", max_sessions=2, max_context=32, max_vocode_frames=group * L + 64)
    e.vocode(codes[: 2 * group * L], cu[: 2 * group + 1], out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(2):
        e.vocode(codes, cu, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 2
    print(f"group {group}: {ms:.1f} ms  {n * L * 132.78e6 / (ms / 1e3) / 1e12:.0f} TFLOP/s  free {torch.cuda.mem_get_info()[0] / 2**30:.0f} GiB", flush=True)
    e.close()
