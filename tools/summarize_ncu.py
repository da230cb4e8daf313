"""Turns ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md>     # per-launch device times of one bench step
  python tools/summarize_ncu.py full <report.ncu-rep> <out.md>       # key metrics of a --set full capture
"""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mn, mv, idc, g = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Grid Size"))
    d = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) > mv:
            d.setdefault(r[idc], {"name": r[kn], "grid": r[g]})[r[mn]] = r[mv]
    agg = OrderedDict()
    total = 0.0
    lines = ["| # | kernel | grid | device time (us) |", "|---|---|---|---|"]
    for i, (k, v) in enumerate(d.items()):
        t = float(v["gpu__time_duration.sum"].replace(",", "")) / 1000.0
        total += t
        short = v["name"].split("(")[0].replace("void ", "")
        agg[short] = agg.get(short, 0.0) + t
        lines.append(f"| {i} | `{short}` | {v['grid']} | {t:.1f} |")
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path})\n\nncu --metrics gpu__time_duration.sum --clock-control none; per-launch times "
                f"are cold-cache and serialised: compare shares, not absolutes.\n\n")
        f.write("## share by kernel\n\n| kernel | total us | share |\n|---|---|---|\n")
        for k, t in sorted(agg.items(), key=lambda x: -x[1]):
            f.write(f"| `{k}` | {t:.1f} | {100 * t / total:.1f}% |\n")
        own = {k: t for k, t in agg.items() if not k.startswith("at::")}
        if len(own) != len(agg):   # torch kernels = the untimed synthetic-data set-up of bench.py (device-side generator)
            tot = sum(own.values())
            f.write("\n## share by kernel, engine kernels only (the `at::` launches above are bench.py's untimed "
                    "synthetic-data set-up)\n\n| kernel | total us | share |\n|---|---|---|\n")
            for k, t in sorted(own.items(), key=lambda x: -x[1]):
                f.write(f"| `{k}` | {t:.1f} | {100 * t / tot:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} us over {len(d)} launches\n\n## launches\n\n" + "\n".join(lines) + "\n")


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
]


def full(path, out):
    txt = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    units = rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"## `{name[:110]}`\n\n| metric | value |\n|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"| {k} | {r[hdr.index(k)]} {units[hdr.index(k)]} |\n")
            stalls = []
            for i, h in enumerate(hdr):
                if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i])))
                    except ValueError:
                        pass
            stalls.sort(key=lambda x: -x[1])
            f.write("| top stalls (warps per issue) | " + ", ".join(f"{k} {v:.2f}" for k, v in stalls[:6]) + " |\n\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
