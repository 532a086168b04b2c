"""Per-instruction stall samples of one kernel from an ncu report captured with --set full --import-source on.
usage: python tools/ncu_hotspots.py <report.ncu-rep> [min_samples] [first last]   (runs `ncu -i ... --page source --csv --print-source sass`)
Prints the launch's duration, the stall-reason totals, the 100-instruction windows holding > 1 % of the samples and every
instruction with at least min_samples samples (index, samples, executions, SASS, top two stall reasons)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(rows[0][1][:90], "| samples", tot, "| instructions", len(data))
agg = {s[6:]: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stall totals:", ", ".join(f"{k} {v}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]))
for a in range(0, len(data), 100):
    s = sum(int(r[ix["# Samples"]]) for r in data[a:a + 100])
    if s > tot * 0.01:
        print(f"  window {a:6d}: {s:5d} samples ({100 * s / tot:4.1f} %)")
for i, r in enumerate(data):
    s = int(r[ix["# Samples"]])
    if (rng and rng[0] <= i <= rng[1]) or (not rng and s >= min_s):
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
        print(f"{i:6d} {s:4d} {r[ix['Instructions Executed']]:>8s}  {r[ix['Source']].strip()[:72]:72s} " + " ".join(f"{k}:{v}" for v, k in top if v > 0))
