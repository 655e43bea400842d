"""Top warp-stall sampling locations of one launch in an .ncu-rep (SASS view).
usage: ncu_top_stalls.py file.ncu-rep [launch-skip] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; skip = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
hdr = rows[1]
si, src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
data = []
for i, r in enumerate(rows[2:]):
    try:
        data.append((float(r[si] or 0), i, r[src].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
for v, i, t in sorted(data, key=lambda x: -x[0])[:top]:
    print("%6.1f%%  #%-5d %s" % (100 * v / tot, i, t[:110]))
