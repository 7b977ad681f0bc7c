"""Executed instructions and stall samples per CUDA source line of one kernel (ncu --import-source on, -lineinfo).
usage: ncu_lines.py report.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
acc = collections.OrderedDict()
fname, hdr, seen_first = None, None, False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if seen_first:
            break                      # next launch
        continue
    if r[0] == "Line No" and "Instructions Executed" in r:
        hdr = r
        ei, si, li = r.index("Instructions Executed"), r.index("Warp Stall Sampling (All Samples)"), 0
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        v, s = int(r[ei]), int(r[si])
    except ValueError:
        continue
    seen_first = True
    key = (fname, r[0])
    a = acc.setdefault(key, [0, 0, r[1]])
    a[0] += v
    a[1] += s
tot = sum(a[0] for a in acc.values()) or 1
stot = sum(a[1] for a in acc.values()) or 1
print("executed %d warp instructions, %d stall samples" % (tot, stot))
for (f, l), (v, s, src) in acc.items():
    if 100.0 * v / tot >= minpct or 100.0 * s / stot >= minpct:
        print("%9d %5.1f%%  stall %5.1f%%  %s:%s  %s" % (v, 100.0 * v / tot, 100.0 * s / stot, f, l, src.strip()[:100]))
