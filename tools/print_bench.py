"""stdin: one bench.py JSON line -> short summary."""
import json, sys
d = json.loads(sys.stdin.read())
r = d.get("roofline", {})
print(round(d["value"], 2), round(d["ms_per_step"], 1), (d.get("e2e") or {}).get("value"), r.get("classes"), d.get("clocks", {}).get("sm_mhz"))
