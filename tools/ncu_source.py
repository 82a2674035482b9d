"""Aggregate the SASS source page of an .ncu-rep: instruction mix and stall samples."""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
ops = collections.Counter()
stalls = collections.Counter()
samples = 0
top = []
for r in rows[hi + 1:]:
    try:
        n = int(r[iE])
    except (ValueError, IndexError):
        continue
    src = r[iS].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    ops[op.split(".")[0]] += n
    tot += n
    try:
        s = int(r[iSamp])
    except ValueError:
        s = 0
    samples += s
    top.append((s, n, src))
    for i, h in stall_cols:
        try:
            stalls[h] += int(r[i])
        except ValueError:
            pass
print("total warp instructions", tot, " samples", samples)
for op, n in ops.most_common(22):
    print("  %-10s %6.2f%%" % (op, 100.0 * n / tot))
print("stall samples:")
for h, n in stalls.most_common(10):
    print("  %-26s %6.2f%%" % (h, 100.0 * n / max(1, sum(stalls.values()))))
if "--top" in sys.argv:
    for s, n, src in sorted(top, reverse=True)[:30]:
        print("  %6d samp %10d exec  %s" % (s, n, src[:100]))
