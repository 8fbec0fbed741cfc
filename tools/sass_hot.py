"""Hot-region view of an ncu source page (SASS): python tools/sass_hot.py prof.ncu-rep <kernel regex> [launch-skip]"""
import collections, csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
hi = his[0]
end = his[1] - 2 if len(his) > 1 else len(rows)
h = rows[hi]
col = {n: h.index(n) for n in h}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[hi + 1:end]:
    try:
        data.append((r[col["Source"]].strip(), int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]),
                     int(r[col["Thread Instructions Executed"]]), [int(r[col[s]] or 0) for s in stalls]))
    except Exception:
        pass
tot = sum(d[2] for d in data); tots = sum(d[1] for d in data)
print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "")
print("sass instr", len(data), "warp instr executed", tot, "samples", tots)
agg = [0] * len(stalls)
for d in data:
    for i, v in enumerate(d[4]): agg[i] += v
print("stall totals:", ", ".join("%s %.1f%%" % (s[6:], 100 * v / max(1, sum(agg))) for s, v in sorted(zip(stalls, agg), key=lambda kv: -kv[1])[:8]))
seg = int(sys.argv[4]) if len(sys.argv) > 4 else 50
for s in range(0, len(data), seg):
    blk = data[s:s + seg]
    ie = sum(d[2] for d in blk); sm = sum(d[1] for d in blk)
    if ie < 0.004 * tot and sm < 0.004 * tots: continue
    ops = collections.Counter()
    for d in blk:
        parts = d[0].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        ops[op.split(".")[0]] += d[2]
    st = [0] * len(stalls)
    for d in blk:
        for i, v in enumerate(d[4]): st[i] += v
    top = ", ".join("%s:%.0f%%" % (k, 100 * v / max(ie, 1)) for k, v in ops.most_common(5))
    tst = ", ".join("%s %.0f%%" % (n[6:], 100 * v / max(1, sum(st))) for n, v in sorted(zip(stalls, st), key=lambda kv: -kv[1])[:3])
    print("%4d-%4d inst %5.1f%% samp %5.1f%% thr %4.1f | %s | %s" % (s, s + seg, 100 * ie / tot, 100 * sm / tots,
          sum(d[3] for d in blk) / max(ie, 1), top, tst))
