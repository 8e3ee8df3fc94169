"""Summarise an ncu report (read on the CPU box): one line per kernel launch with the metrics the roofline uses."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel", None), ("gpu__time_duration.sum", "us", 1), ("launch__registers_per_thread", "regs", 0),
        ("launch__grid_size", "grid", 0), ("dram__bytes_read.sum", "dram_rd", 2), ("dram__bytes_write.sum", "dram_wr", 2),
        ("lts__t_bytes.sum", "l2_bytes", 2), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1),
        ("smsp__inst_executed.sum", "warp_inst", 0),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
        ("lts__t_sector_hit_rate.pct", "l2hit%", 1)]
print("| " + " | ".join(c[1] for c in cols) + " |")
print("|" + "---|" * len(cols))
for r in rows[2:]:
    out = []
    for name, label, nd in cols:
        if name not in idx:
            out.append("-")
            continue
        v = r[idx[name]]
        if nd is None:
            v = v.split("(")[0].replace("void ", "").strip()[:48]
        else:
            try:
                f = float(v.replace(",", ""))
                u = units[idx[name]]
                if label in ("dram_rd", "dram_wr", "l2_bytes"):
                    mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    v = "%.1f MB" % (f * mult)
                elif label == "us":
                    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
                    v = "%.1f" % (f * mult)
                else:
                    v = ("%." + str(nd) + "f") % f
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
