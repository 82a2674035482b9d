"""Aggregate an .ncu-rep per CUDA source line: warp instructions executed and stall samples (top N lines)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = ""
agg = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r and r[0].isdigit():
        try:
            agg.append((int(r[iE]), int(r[iS]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
tot_i = sum(a[0] for a in agg) or 1
tot_s = sum(a[1] for a in agg) or 1
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
byfile = {}
for a in agg:
    byfile.setdefault(a[2], [0, 0])
    byfile[a[2]][0] += a[0]; byfile[a[2]][1] += a[1]
for f, (i, s) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("  %-28s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * i / tot_i, 100.0 * s / tot_s))
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% samp  %s:%d  %s" % (100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s, a[2], a[3], a[4][:90]))
