"""Developer tool: instruction / sample share per code region of phase1_kernel from an ncu report (--import-source on).
Usage: python tools/ncu_regions.py gpurun_out/prof.ncu-rep [launch-index]   (regions are found by marker strings in the source)"""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = open(os.path.join(ROOT, "cpuvox_b200", "csrc", "raybuffer_kernels.cu")).read().split("\n")
def line_of(marker, start=0):
    for i in range(start, len(src)):
        if marker in src[i]: return i + 1
    raise SystemExit("marker not found: " + marker)
marks = [("math helpers", "struct F3 {"), ("dda", "struct Dda {"), ("clip helpers", "cross2(float ax"), ("ray setup", "struct RaySetup {"),
         ("mask helpers", "uint32_t mask_from(int a)"), ("span_would_write", "bool span_would_write("), ("reduce_horizon", "void reduce_pixel_horizon("),
         ("misc", "struct Acc {"), ("prologue", "phase1_kernel(const __grid_constant__"), ("walk+hdr", "EMU_STAT(0); // batches"), ("hull", "int hullMin = INT_MIN / 2"),
         ("select", "EMU_STAT(1); // select"), ("renarrow", "EMU_STAT(2); // renarrows"), ("column setup", "const uint32_t* colColors ="),
         ("round form", "EMU_STAT(4); // rounds formed"), ("FAST geometry", "// ---- one boundary per lane: record k"), ("general geometry", "uint32_t el = 0u;"),
         ("round hot test", "roundHotS = GBALLOT("), ("resolve", "// ---- resolve this column (pass)"), ("commit ctl", "uint32_t candS = GBALLOT(sideOk)"),
         ("commit horizon+pixels", "reduce_pixel_horizon(rw, bMin, bMax); // :507"), ("commit mark", "mark_seen<G>(rw.seen"), ("epilogue", "if (pendY >= 0) row[pendY] = pendColor;\n"), ("phase2+rest", "// ---- Phase 2 ---")]
bounds = []
for name, m in marks:
    try: bounds.append((name, line_of(m)))
    except SystemExit: pass
bounds.sort(key=lambda x: x[1])
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
blocks, cur = [], []
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "File Path":
        if cur: blocks.append(cur)
        cur = []
    cur.append(row)
if cur: blocks.append(cur)
b = [b for b in blocks if b[0][1].endswith(".cu")][which]
hdr = next(r for r in b if r and r[0] == "Line No"); ix = {h: i for i, h in enumerate(hdr)}
rows = [r for r in b if r and r[0].isdigit()]
ti = sum(int(r[ix["Instructions Executed"]]) for r in rows); ts = sum(int(r[ix["# Samples"]]) for r in rows)
print(b[1][1][:110]); print(f"total warp instructions {ti}  samples {ts}")
for k, (name, lo) in enumerate(bounds):
    hi = bounds[k + 1][1] - 1 if k + 1 < len(bounds) else 10**9
    i = sum(int(r[ix["Instructions Executed"]]) for r in rows if lo <= int(r[0]) <= hi)
    s = sum(int(r[ix["# Samples"]]) for r in rows if lo <= int(r[0]) <= hi)
    if i: print(f"{name:24s} lines {lo:4d}-{hi if hi < 10**9 else 0:4d}  inst {100 * i / ti:5.1f}%  samples {100 * s / ts:5.1f}%")
