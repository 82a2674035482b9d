"""Bin the SASS of an .ncu-rep by how often each instruction runs per warp: separates the leapfrog loop from the
once-per-iteration transition / refresh code.  usage: ncu_bins.py rep n_warps"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
nw = float(sys.argv[2])
bins = collections.defaultdict(lambda: [0, 0, 0])
for r in rows[hi + 1:]:
    try:
        n = int(r[iE]); s = int(r[iSamp])
    except (ValueError, IndexError):
        continue
    key = round(n / nw, 1)
    b = bins[key]; b[0] += 1; b[1] += n; b[2] += s
tot = sum(b[1] for b in bins.values()); ts = sum(b[2] for b in bins.values())
print("exec/warp  #sass  share_of_instr  share_of_samples")
for k in sorted(bins, key=lambda k: -bins[k][1])[:25]:
    b = bins[k]
    print("%9.1f %6d %8.2f%% %8.2f%%" % (k, b[0], 100.0 * b[1] / tot, 100.0 * b[2] / max(ts, 1)))
