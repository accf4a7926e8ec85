"""Top stall sites of one kernel from `ncu -i REP --page source --csv --kernel-name regex:NAME` output (SASS view).

    python tools/ncu_src_top.py cols16_src.csv [N]
Prints the N instructions with the most stall samples, the dominant stall reason of each, and totals by reason
and by opcode class -- the numbers quoted in profiles/README.md.
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and not h.endswith("_not_issued")]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[ix["# Samples"]] != "# Samples"]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
by_reason, by_op = Counter(), Counter()
items = []
for k, r in enumerate(body):
    s = int(r[ix["# Samples"]] or 0)
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    by_op[op.split(".")[0]] += s
    reasons = {c: int(r[ix[c]] or 0) for c in stall_cols}
    for c, v in reasons.items():
        by_reason[c] += v
    items.append((s, k, r[ix["Source"]][:70], max(reasons, key=reasons.get) if s else "", int(r[ix["Instructions Executed"]] or 0)))
print(f"total samples {tot}, instructions {len(body)}")
print("by reason:", ", ".join(f"{c[6:]} {100 * v / max(1, sum(by_reason.values())):.1f}%" for c, v in by_reason.most_common(8)))
print("by opcode:", ", ".join(f"{o} {100 * v / max(1, tot):.1f}%" for o, v in by_op.most_common(10)))
for s, k, src, why, ex in sorted(items, reverse=True)[:n]:
    print(f"{100 * s / max(1, tot):5.1f}%  line {k:5d}  exec {ex:8d}  {why[6:]:22s} {src}")
