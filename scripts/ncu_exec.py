"""Executed-instruction profile of one kernel from an .ncu-rep (SASS view).  usage: ncu_exec.py report kernel_regex [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[start]
ei, src, si = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for k, r in enumerate(rows[start + 1:]):
    if r == hdr:
        break
    try:
        data.append((int(r[ei]), int(r[si]), k, r[src]))
    except (ValueError, IndexError):
        pass
tot = sum(v for v, _, _, _ in data) or 1
print("total executed", tot, "instructions", len(data))
# contiguous regions with similar counts
for v, st, k, s in data:
    if v * 200 >= tot * 1 or st > 15:
        print("%9d %5.1f%% st=%4d #%-5d %s" % (v, 100.0 * v / tot, st, k, s[:110]))
