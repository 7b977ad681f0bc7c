"""Per-kernel table from an ncu --set full report: duration, DRAM traffic, L2 sectors, FP32 pipe, IPC, occupancy.
usage: ncu_summary.py report.ncu-rep [kernel_regex]"""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
rx = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError, IndexError):
        return d
units = rows[1]
def unit(k):
    return units[ix[k]] if k in ix else ""
print("| kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 sectors M | FP32 pipe % | warp instr M | IPC / SM | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    short = re.sub(r"^.*?::", "", re.sub(r"\(.*$", "", name)).replace("unnamed>::", "").lstrip("<")
    if rx and not rx.search(name):
        continue
    us = g(r, "gpu__time_duration.sum")
    if unit("gpu__time_duration.sum") == "ns": us /= 1e3
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    for k, v in (("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr")):
        u = unit(k)
        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if v == "rd": rd *= f
        else: wr *= f
    fp32 = g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", g(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"))
    print("| %s | %d x %d | %d | %.2f | %.2f | %.2f | %.0f | %.2f | %.1f | %.2f | %.2f | %.1f |" % (
        short[:44], g(r, "launch__grid_size"), g(r, "launch__block_size"), g(r, "launch__registers_per_thread"), us, rd, wr,
        (rd + wr) / us * 1e3 if us else 0, g(r, "lts__t_sectors.sum") / 1e6, fp32,
        g(r, "smsp__inst_executed.sum") / 1e6, g(r, "sm__inst_executed.avg.per_cycle_active"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")))
