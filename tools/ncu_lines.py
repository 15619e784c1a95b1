"""Developer tool: per-source-line hot spots of an ncu report (needs -lineinfo + --import-source on).
Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top] [launch-index]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur = [], []
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "File Path":
        if cur:
            blocks.append(cur)
        cur = []
    cur.append(row)
if cur:
    blocks.append(cur)
# keep the blocks of our .cu file
mine = [b for b in blocks if b[0][1].endswith(".cu")]
b = mine[which]
hdr = next(r for r in b if r and r[0] == "Line No")
ix = {h: i for i, h in enumerate(hdr)}
rows = [r for r in b if r and r[0].isdigit()]
tot_s = sum(int(r[ix["# Samples"]]) for r in rows) or 1
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in rows) or 1
print(f"{b[0][1]}  {b[1][1]}  samples={tot_s} inst={tot_i}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
rows.sort(key=lambda r: -int(r[ix["# Samples"]]))
for r in rows[:top]:
    s = int(r[ix["# Samples"]])
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{int(r[0]):5d} {100*s/tot_s:5.1f}%smp {100*int(r[ix['Instructions Executed']])/tot_i:5.1f}%inst  "
          f"{' '.join(f'{n}:{c}' for c, n in st if c):40s} | {r[1].strip()[:110]}")
