"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md): tcgen05 MMAs (UTCHMMA /
UTCQMMA), TMEM loads (LDTM), tensor-memory barriers (UTCBAR), TMA tensor loads (UTMALDG), bulk copies (UBLKCP), legacy warp
MMAs (HMMA), DSMEM / cluster instructions.

    python scripts/sass_summary.py > profiles/sass_summary.md          # needs cuobjdump (CUDA toolkit), no GPU"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "llmvox_b200", "libllmvox_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "UTMAPF", "HMMA", "SYNCS", "UCGABAR", "MUFU", "FFMA"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        counts[cur]["_lines"] = 0
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_lines"] += 1
        if op in MNEMONICS:
            counts[cur][op] += 1
        if ".CLUSTER" in m.group(1) or "MAPA" in m.group(1) or op in ("UCGABAR_ARV", "UCGABAR_WAIT"):
            counts[cur]["cluster/DSMEM"] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS summary of `llmvox_b200/libllmvox_b200.so` (sm_100a)\n")
print("`cuobjdump -sass llmvox_b200/libllmvox_b200.so`, counted by `scripts/sass_summary.py` (no GPU needed; regenerate after every")
print("kernel change).  UTCHMMA = `tcgen05.mma.kind::f16`, LDTM = `tcgen05.ld`, UTCBAR = `tcgen05.commit`, UTMALDG = TMA tensor load,")
print("UBLKCP = `cp.async.bulk`, HMMA = legacy warp MMA (`mma.sync`), SYNCS = mbarrier ops.\n")
cols = MNEMONICS + ["cluster/DSMEM"]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---:|" + "---:|" * len(cols))
for (k, c), nm in zip(counts.items(), names):
    if not any(c[x] for x in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "cluster/DSMEM")):
        continue
    nm = re.sub(r"\(.*", "", nm).replace("void ", "")
    print(f"| `{nm}` | {c['_lines']} | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
others = [re.sub(r"\(.*", "", nm).replace("void ", "") for (k, c), nm in zip(counts.items(), names)
          if not any(c[x] for x in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UBLKCP", "HMMA", "cluster/DSMEM"))]
print("\nKernels without tensor-core / TMA / cluster instructions (glue, FMA-pipe parity mode): " + ", ".join(f"`{x}`" for x in sorted(set(others))))
