"""Condense an `ncu --page raw --csv` capture of one pass (tools/gpu_round.sh) into a per-launch table:
   python tools/ncu_summarize.py gpurun_out/<tag>_step_sections.csv [ops_profile.json [traffic.json]] > profiles/<tag>_step_summary.md
With a third argument also writes {launch name: DRAM bytes read + written per launch} (what bench.py reports as
roofline.traffic)."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
col = {c: i for i, c in enumerate(h)}
data = [r for r in rows[hi + 2:] if len(r) == len(h)]
names = None
if len(sys.argv) > 2:
    names = [o["name"] for o in json.load(open(sys.argv[2]))]


def f(r, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except (ValueError, KeyError):
        return float("nan")


print("| # | launch | kernel | grid | regs | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | dram % | L2 % | tensor % (active) | sm % | warps act % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
tot = 0.0
tr = tw = 0.0
traffic = {}
for i, r in enumerate(data):
    k = r[col["Kernel Name"]].split("(")[0].replace("xrseg::", "").replace("void ", "")
    us = f(r, "gpu__time_duration.sum") / 1e3
    rd, wr = f(r, "dram__bytes_read.sum") / 1e6, f(r, "dram__bytes_write.sum") / 1e6
    tot += us
    tr += rd
    tw += wr
    nm = names[i] if names and i < len(names) else ""
    if nm and nm not in traffic:
        traffic[nm] = {"dram_bytes": (rd + wr) * 1e6, "us_under_ncu": us, "kernel": k,
                       "tensor_pct": f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
    print(f"| {i} | {nm} | {k[:28]} | {r[col['Grid Size']]} | {int(f(r, 'launch__registers_per_thread'))} | {us:.1f} | {rd:.1f} | {wr:.1f} | "
          f"{(rd + wr) / us * 1e3:.0f} | {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} |")
print(f"\nlaunches {len(data)}; sum of durations {tot:.0f} us; DRAM read {tr:.0f} MB + write {tw:.0f} MB = {tr + tw:.0f} MB per pass")
if len(sys.argv) > 3:
    json.dump({"source": sys.argv[1], "batch": 64, "launches": traffic}, open(sys.argv[3], "w"), indent=0)
