"""Per captured launch of an .ncu-rep (source page): warp-stall samples and executed instructions by opcode class,
and the hottest instructions."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for n, hi in enumerate(heads):
    end = heads[n + 1] - 1 if n + 1 < len(heads) else len(rows)
    name = rows[hi - 1][1][:140] if hi and len(rows[hi - 1]) > 1 else (rows[hi - 1][0][:140] if hi else "")
    h = rows[hi]
    isrc, ist, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    data = []
    for i, r in enumerate(rows[hi + 1:end]):
        if len(r) > ist and r[ist].isdigit():
            data.append((int(r[ist]), int(r[iex]) if r[iex].isdigit() else 0, r[isrc].strip(), i))
    tot = max(1, sum(d[0] for d in data))
    totex = max(1, sum(d[1] for d in data))
    print("==== launch %d: %s" % (n, name))
    print("samples %d  warp-instructions %d  sass lines %d" % (tot, totex, len(data)))
    cls = defaultdict(lambda: [0, 0])
    for s, e, src, i in data:
        if not src:
            continue
        op = src.split()[0] if not src.startswith("@") else src.split()[1]
        op = op.split(".")[0]
        cls[op][0] += s
        cls[op][1] += e
    print("-- by opcode (share of stall samples, share of executed instructions)")
    for op, (s, e) in sorted(cls.items(), key=lambda x: -x[1][1])[:22]:
        print("%-10s %6.2f%% %6.2f%%" % (op, 100.0 * s / tot, 100.0 * e / totex))
    print("-- hottest instructions")
    for s, e, src, i in sorted(data, reverse=True)[:top]:
        print("%6.2f%% ex=%10d #%4d %s" % (100.0 * s / tot, e, i, src[:100]))
