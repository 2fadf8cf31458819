"""Development aid: one-line digest of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d.get("dtype", "")[:40], "| ms", round(d["ms_per_step"], 3), "rows/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      "roof", round(d.get("roofline", {}).get("frac", 0), 4), {k: v for k, v in list(d.get("op_ms_per_step", {}).items())[:9]})
