#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv` dump: instruction mix by opcode and the top stall instructions.
    ncu -i prof.ncu-rep --page source --csv -c 1 > src.csv ; python tools/ncu_source_hot.py src.csv [view=sass|source]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] in ("Address", "Kernel Name")), len(rows))
hdr, body = rows[start], rows[start + 1:end]
ci = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ci[k]])
    except Exception:
        return 0.0


tot, tot_i = sum(f(r, "# Samples") for r in body), sum(f(r, "Instructions Executed") for r in body)
print("samples %d  warp instructions %.3e" % (tot, tot_i))
op, ops = collections.Counter(), collections.Counter()
for r in body:
    s = r[ci["Source"]].strip().split()
    if not s:
        continue
    o = (s[1] if s[0].startswith("@") and len(s) > 1 else s[0]).split(".")[0]
    op[o] += f(r, "Instructions Executed")
    ops[o] += f(r, "# Samples")
for o, c in op.most_common(24):
    print("%-8s instr %5.1f%%  samples %5.1f%%" % (o, 100 * c / tot_i, 100 * ops[o] / tot))
print("--- top stall instructions")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    st = {k: f(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5.2f%% %-60s %s  L2sec %s/%s" % (100 * f(r, "# Samples") / tot, r[ci["Source"]].strip()[:60],
                                              " ".join("%s=%d" % (k[6:], v) for k, v in top),
                                              r[ci["L2 Theoretical Sectors Global"]], r[ci["L2 Theoretical Sectors Global Ideal"]]))
