"""Turn the ncu artefacts a profiles/run_*.sh call left in gpurun_out/ into the markdown tables of profiles/summary_*.md.

    python profiles/summarize.py launches gpurun_out/launches_r1e.csv [first_row last_row]
    python profiles/summarize.py full gpurun_out/prof_gemm_r1e.ncu-rep [...]

`launches`: per-kernel totals of a `--metrics gpu__time_duration.sum` launch list (cold-cache, serialised times:
compare SHARES with bench.py, not absolutes). `full`: key columns of `--set full` captures, one row per launch.
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("fx::", "")


def launches(path, lo=None, hi=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = [r for r in rows if r[im] == "gpu__time_duration.sum"]
    if lo is not None:
        rows = rows[lo:hi]
    agg = OrderedDict()
    for r in rows:
        v = float(r[iv].replace(",", ""))
        us = {"ns": v / 1e3, "us": v, "usecond": v, "nsecond": v / 1e3, "ms": v * 1e3, "msecond": v * 1e3}[r[iu]]
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.1f} ms serialised\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us / 1e3:.2f} | {100 * us / tot:.1f}% | {us / n:.1f} |")
    fam = {"gemm": 0.0, "fmha": 0.0, "other": 0.0}
    for k, (n, us) in agg.items():
        fam["gemm" if "gemm" in k else "fmha" if "fmha" in k else "other"] += us
    print("\nFamilies: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in fam.items()))


COLS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 thr %"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"), ("launch__grid_size", "grid")]


def full(paths):
    print("| capture | kernel | " + " | ".join(c[1] for c in COLS) + " |\n|---|---|" + "---:|" * len(COLS))
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, rows = rows[0], rows[1], rows[2:]

        def col(name):
            for i, h in enumerate(hdr):
                if h == name or h.endswith("." + name):
                    return i
            return None
        for r in rows:
            cells = []
            for name, _ in COLS:
                i = col(name)
                cells.append("-" if i is None else f"{r[i]} {units[i]}".strip())
            print(f"| {path.split('/')[-1].replace('.ncu-rep', '')} | `{short(r[hdr.index('Kernel Name')])}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        a = sys.argv[3:]
        launches(sys.argv[2], *(int(x) for x in a))
    else:
        full(sys.argv[2:])
