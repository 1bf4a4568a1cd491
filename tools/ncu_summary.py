"""Compact per-kernel table from an .ncu-rep (tools, not product): usage ncu_summary.py <rep> > profiles/x.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__inst_executed.sum", "warp inst"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem")]
print(f"# ncu summary of `{rep.split('/')[-1]}` (--set full --clock-control none; cold-cache, serialised launches)\n")
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("apk::", "")[:70]
    vals = []
    for key, _ in cols:
        if key in ix:
            v = r[ix[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[ix[key]]}".strip())
        else:
            vals.append("-")
    print(f"| `{name}` | " + " | ".join(vals) + " |")
