#!/bin/bash
# per-kernel time of one 80-frame decode (ncu launch list, serialised) + the un-profiled wall time
python tools/prof_decoder.py 80
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/dec_launches.csv python tools/prof_decoder.py 80 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/dec_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.defaultdict(float); c = collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    name = r[ki].split("(")[0].split("::")[-1][:50]
    t[name] += v; c[name] += 1
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:14]:
    print(f"{k:52s} {c[k]/5:7.1f} launches/decode {v/5:8.2f} ms/decode")
PY
