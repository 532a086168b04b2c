"""Top SASS instructions / CUDA source lines by stall samples from an ncu report:
   python tools/ncu_hot.py <rep> [kernel-index] [cuda|sass] [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"Kernel Name"') or line.startswith('"File Path"') and view == "cuda" and False:
        if cur:
            blocks.append(cur)
        cur = []
    cur.append(line)
if cur:
    blocks.append(cur)
blocks = [b for b in blocks if b and b[0].startswith('"Kernel Name"')]
b = blocks[kidx]
print(b[0][:120])
rows = list(csv.reader(io.StringIO("\n".join(b[1:]))))
hdr = None
data = []
for r in rows:
    if "Source" in r and "# Samples" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {c: i for i, c in enumerate(hdr)}
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
s_i = ci["# Samples"]
src_i = ci["Source"]
tot = sum(int(r[s_i] or 0) for r in data)
exe = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "warp instructions executed", exe)
agg = {c: sum(int(r[ci[c]] or 0) for r in data) for c in stall_cols}
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
data.sort(key=lambda r: -int(r[s_i] or 0))
for r in data[:top]:
    st = sorted(((c, int(r[ci[c]] or 0)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
    key = r[ci["Line No"]] + ": " if "Line No" in ci else r[ci["Address"]][-5:] + " "
    print(f"{int(r[s_i]):7d} {100 * int(r[s_i]) / max(tot, 1):5.1f}%  exec {r[ci['Instructions Executed']]:>9s}  {key}{r[src_i].strip()[:110]}   {st}")
