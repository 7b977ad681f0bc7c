"""Top stall sites of one kernel from an .ncu-rep (SASS view).  usage: ncu_hot.py report.ncu-rep kernel_regex [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[start]
si, src = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
data = []
for k, r in enumerate(rows[start + 1:]):
    if r == hdr:
        break            # next launch of the kernel
    try:
        data.append((int(r[si]), k, r[src]))
    except (ValueError, IndexError):
        pass
tot = sum(v for v, _, _ in data) or 1
print("total samples", tot, "instructions", len(data))
for v, k, s in sorted(data, key=lambda x: -x[0])[:n]:
    print("%6d %5.1f%%  #%-5d %s" % (v, 100.0 * v / tot, k, s[:140]))
