"""Per-kernel table (markdown + json) from an ncu --set full report: one row per distinct kernel (first launch).
usage: ncu_table.py report.ncu-rep out_prefix [kernel_regex]"""
import csv, io, json, re, subprocess, sys
rep, outp = sys.argv[1], sys.argv[2]
rx = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError, IndexError):
        return d
def scaled(r, k):
    v = g(r, k)
    u = units[ix[k]] if k in ix else ""
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
STALLS = ["long_scoreboard", "short_scoreboard", "lg_throttle", "mio_throttle", "math_pipe_throttle", "barrier", "wait",
          "membar", "branch_resolving", "no_instruction", "not_selected", "dispatch_stall", "drain", "imc_miss", "tex_throttle", "sleeping"]
table, seen = {}, set()
md = ["| kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | L2 sectors M | FP32 pipe % | warp instr M | IPC / SM | warps active % | top stall |",
      "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    if rx and not rx.search(name):
        continue
    short = re.sub(r"\(.*$", "", name)
    short = re.sub(r"^.*::", "", short)
    if short in seen:
        continue
    seen.add(short)
    us = scaled(r, "gpu__time_duration.sum")
    rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
    fp32 = g(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", g(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"))
    stalls = sorted(((g(r, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s), s) for s in STALLS), reverse=True)
    top = "%s %.1f" % (stalls[0][1], stalls[0][0])
    base = re.sub(r"<.*$", "", short)
    entry = {"name": short, "grid": int(g(r, "launch__grid_size")), "block": int(g(r, "launch__block_size")),
             "regs": int(g(r, "launch__registers_per_thread")), "us": us, "dram_read_bytes": rd, "dram_write_bytes": wr,
             "dram_bytes": rd + wr, "l2_sectors": g(r, "lts__t_sectors.sum"), "fp32_pipe_pct": fp32,
             "warp_instructions": g(r, "smsp__inst_executed.sum"), "ipc_per_sm": g(r, "sm__inst_executed.avg.per_cycle_active"),
             "warps_active_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
             "limiter": "top stall %s inst/issue; DRAM %.2f TB/s; IPC %.2f" % (top, (rd + wr) / us / 1e6 if us else 0, g(r, "sm__inst_executed.avg.per_cycle_active"))}
    table.setdefault(base, entry)
    table[short] = entry
    md.append("| %s | %d x %d | %d | %.2f | %.2f | %.2f | %.0f | %.2f | %.1f | %.2f | %.2f | %.1f | %s |" % (
        short[:60], entry["grid"], entry["block"], entry["regs"], us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3 if us else 0,
        entry["l2_sectors"] / 1e6, fp32, entry["warp_instructions"] / 1e6, entry["ipc_per_sm"], entry["warps_active_pct"], top))
json.dump({"source": rep.split("/")[-1], "kernels": table}, open(outp + ".json", "w"), indent=1)
open(outp + ".md", "w").write("\n".join(md) + "\n")
print("\n".join(md))
