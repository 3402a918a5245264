"""Summarise the source page of an ncu report (ncu -i rep --page source --csv > file.csv): stall samples per stall
reason, and the instructions / SASS regions where they fall. usage: python tools/ncu_source_hot.py file.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = rows[2:]
tot = Counter()
n_samples = 0
per = []
for r in data:
    if len(r) < len(hdr):
        continue
    s = int(r[col["# Samples"]] or 0)
    n_samples += s
    for st in stalls:
        v = int(r[col[st]] or 0)
        tot[st] += v
    per.append((s, r[col["Address"]], r[col["Source"]], int(r[col["Instructions Executed"]] or 0),
                {st: int(r[col[st]] or 0) for st in stalls if int(r[col[st]] or 0)}))
print("total samples", n_samples, "instructions", len(per), "executed", sum(p[3] for p in per))
for st, v in tot.most_common():
    if v:
        print(f"  {st:28s} {v:8d} {100.0 * v / max(n_samples, 1):5.1f}%")
# opcode classes
ops = Counter(); opsamp = Counter()
for s, a, src, ex, st in per:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += ex; opsamp[op] += s
print("opcode: executed, samples")
for op, v in ops.most_common(25):
    print(f"  {op:10s} {v:9d} {opsamp[op]:8d}")
print("hottest instructions:")
for s, a, src, ex, st in sorted(per, key=lambda x: -x[0])[:top]:
    print(f"  {s:6d} {ex:8d} {src[:70]:70s} {st}")
