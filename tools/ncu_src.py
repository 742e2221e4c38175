"""Summarise `ncu -i X.ncu-rep --page source --csv`: executed warp instructions and stall samples per
SASS block (blocks are cut where the executed count changes by more than 20 %)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[ia], r[isrc], int(r[iex] or 0), int(r[ismp] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[2] for d in data); tots = sum(d[3] for d in data)
print("total warp instr", tot, "samples", tots)
blocks = []; cur = []
for d in data:
    if cur and (abs(d[2] - cur[-1][2]) > 0.2 * max(d[2], cur[-1][2], 1)):
        blocks.append(cur); cur = []
    cur.append(d)
if cur: blocks.append(cur)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
for b in blocks:
    s = sum(x[2] for x in b); sm = sum(x[3] for x in b)
    if 100.0 * s / tot < thr and 100.0 * sm / max(tots, 1) < thr: continue
    ops = {}
    for x in b:
        op = x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:8])
    print(f"{b[0][0][-5:]}-{b[-1][0][-5:]} n={len(b):3d} exec/instr={b[0][2]:>10d} instr%={100.0*s/tot:5.1f} stall%={100.0*sm/max(tots,1):5.1f}  {top}")
