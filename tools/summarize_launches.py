"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list: time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [l for l in open(path, errors="replace") if l.startswith('"')]
rd = csv.reader(rows)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if len(r) <= iv:
        continue
    name = re.sub(r"^(void )?(oob::)?", "", r[ik]).split("<")[0].split("(")[0]
    v = float(r[iv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"captured {sum(cnt.values())} launches, total {total:.2f} ms; per-launch times are cold-cache + serialised: compare SHARES\n")
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f"{k:45s} n={cnt[k]:5d} total={tot[k]:9.3f} ms  {100 * tot[k] / total:5.1f}%  avg={1e3 * tot[k] / cnt[k]:8.1f} us")
