"""Executed warp instructions of one kernel by SASS opcode (first launch in the report).  usage: ncu_ops.py report kernel_regex"""
import csv, subprocess, sys, io, collections, re
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[start]
ei, src, si = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
acc, st = collections.Counter(), collections.Counter()
for r in rows[start + 1:]:
    if r == hdr:
        break
    try:
        v, s = int(r[ei]), int(r[si])
    except (ValueError, IndexError):
        continue
    t = r[src].split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED", "SHFL")) else op.split(".")[0]
    acc[op] += v
    st[op] += s
tot, stot = sum(acc.values()) or 1, sum(st.values()) or 1
print("total executed %d, stall samples %d" % (tot, stot))
for op, v in acc.most_common(30):
    print("%-12s %9d %5.1f%%   stall %5.1f%%" % (op, v, 100.0 * v / tot, 100.0 * st[op] / stot))
