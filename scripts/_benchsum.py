"""One-line digest of bench.py output logs: python scripts/_benchsum.py LOG..."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        o = d.get("other_precision") or {}
        s = d.get("streams256") or {}
        v = d.get("vocoder_bulk") or {}
        print(f"{path}: {d['config'].get('precision')} value {d['value']:.0f} ms/step {d['ms_per_step']:.2f} e2e {d['e2e']['value']:.0f} "
              f"roofline {r.get('kernel')} {r.get('frac', 0):.3f} | other {o.get('precision')} {o.get('value', 0):.0f} | "
              f"256: {s.get('value', 0):.0f} e2e {s.get('e2e', 0):.0f} {s.get('decode_path', '')} | voc {v.get('tflops', 0):.0f} TF")
