"""Developer tool: turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.
Usage: python tools/summarize_profiles.py <tag>   (reads gpurun_out/launches.csv and gpurun_out/prof_phase1_4k.ncu-rep)"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)


def launches():
    p = os.path.join(ROOT, "gpurun_out", "launches.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == "ID")
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if r[0] == "ID" or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        us = v / 1e3 if u.startswith("ns") else v if u.startswith("us") else v * 1e3
        k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += us
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-1080p`\n\n")
        f.write(f"{n} launches captured (cold-cache, serialised: compare shares, not absolutes). phase1_kernel<32,1,0> is the counters build\n"
                "bench.py runs once per pose before the timed region to obtain the algorithmic byte counts.\n\n")
        f.write("| kernel | launches | total ms | share | avg us | grid | block |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f}% | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} |\n")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warps_active_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
]


def full(rep, kernel="phase1", traffic_file="phase1_traffic.json"):
    p = os.path.join(ROOT, "gpurun_out", rep + ".ncu-rep")
    if not os.path.exists(p):
        return
    txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    with open(os.path.join(out, f"{tag}_{rep}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kernel} ({tag}, {rep})\n\n")
        f.write("Command: `python tools/one_frame.py --res 3840x2160 --poses 30,59 --reps 2` (mill 1024^3; the captured launches are pose 59, the heaviest class of the path)\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
        for w in WANT + stall:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"| `{w}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
    i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    tr = sum(tobytes(r[i_r], units[i_r]) + tobytes(r[i_w], units[i_w]) for r in data) / len(data)
    tp = os.path.join(out, traffic_file)
    t = json.load(open(tp)) if os.path.exists(tp) else {}
    t["3840"] = tr
    t["_source"] = f"profiles/{tag}_{rep}.md (pose 59, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(t, open(tp, "w"), indent=1)
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), p, "45"], capture_output=True, text=True).stdout
    with open(os.path.join(out, f"{tag}_{rep}_lines.txt"), "w") as f:
        f.write(lines)


def instructions():
    """profiles/phase1_inst.json: mean warp instructions per product Phase-1 launch over the 60 poses (gpurun_out/inst.csv)."""
    p = os.path.join(ROOT, "gpurun_out", "inst.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = next(r for r in rows if r[0] == "ID")
    ix = {h: i for i, h in enumerate(hdr)}
    vals = [float(r[ix["Metric Value"]].replace(",", "")) for r in rows
            if r[0] != "ID" and r[ix["Metric Name"]] == "smsp__inst_executed.sum" and "phase1_kernel<32, 0, 0" in r[ix["Kernel Name"]].replace("(int)", "").replace("(bool)", "")]
    vals = vals[-60:]   # the timed step's 60 product launches come last (warm-up and exclusive passes before them use the same kernel)
    if not vals:
        return
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"3840": sum(vals) / len(vals), "_per_pose_min_max": [min(vals), max(vals)],
               "_source": f"ncu --metrics smsp__inst_executed.sum over the last 60 product Phase-1 launches of `bench.py --steps 1 --warmup 1 --no-extras` at 3840x2160 "
                          f"(tools/gpu_round.sh, {tag}, kernel of commit {head})"},
              open(os.path.join(out, "phase1_inst.json"), "w"), indent=1)


launches()
instructions()
full("prof_phase1_4k")
full("prof_phase2_4k", "phase2", "phase2_traffic.json")
